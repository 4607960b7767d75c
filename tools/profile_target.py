"""Small fixed workload for ncu: reorient, NENV envs, a few env-steps of 5 substeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dexterity_b200 import backend
if os.environ.get("DEX_LIB"): backend.LIB_PATH = os.environ["DEX_LIB"]
from dexterity_b200.model import scenes
from tests import statelib

nenv = int(os.environ.get("NENV", 16384))
steps = int(os.environ.get("STEPS", 3))
compiled = scenes.load_compiled("reorient_shadow")
model = backend.PhysicsModel(compiled, 0)
states = statelib.rollout_states(compiled, seed=123, nstates=256, every=9, episodes=16)
b = backend.PhysicsBatch(model, nenv, torch.float32)
idx = np.arange(nenv) % 256
statelib.upload(b, {k: s[idx] for k, s in states.items()})
lo = torch.as_tensor(compiled.ACT_CTRLRANGE[:, 0], dtype=torch.float32, device="cuda")
hi = torch.as_tensor(compiled.ACT_CTRLRANGE[:, 1], dtype=torch.float32, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
for k in range(steps):
    b.ctrl.copy_(lo + (hi - lo) * torch.rand(nenv, compiled.nu, device="cuda", generator=g))
    b.step(5)
torch.cuda.synchronize()
print("done", b.ncon.float().mean().item())
