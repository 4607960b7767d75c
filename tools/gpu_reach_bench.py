"""BASELINE.json configs[0..1]: reach (Shadow hand, state obs, random actions): one env on one host core through the
FP64 CPU restatement, and 4096 envs on one B200 through the public batched API."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dexterity_b200 import envs
from dexterity_b200.model import scenes
from oracle import oracle as orc
from tests import statelib

c = scenes.load_compiled("reach_shadow")
om = orc.OracleModel(c)
rng = np.random.RandomState(12345)
q = statelib.reorient_reset(c, rng)[None].copy(); v = np.zeros((1, c.nv)); w = np.zeros((1, c.nv))
lo, hi = c.ACT_CTRLRANGE[:, 0], c.ACT_CTRLRANGE[:, 1]
for _ in range(100):
    om.step_batch(q, v, rng.uniform(lo, hi, size=(1, c.nu)), w, nsub=1, nthreads=1)
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 5.0:
    om.step_batch(q, v, rng.uniform(lo, hi, size=(1, c.nu)), w, nsub=1, nthreads=1); n += 1
print("config 0: reach, 1 env, 1 host core, FP64 restatement (not MuJoCo): %.0f env-steps/s (1 mj_step each)" % (n / (time.perf_counter() - t0)))

N = 4096
env = envs.load("reach", "state_dense", seed=12345, num_envs=N, hand=os.environ.get("HAND", "shadow_hand_e"))
spec = env.action_spec()
lo_t = torch.as_tensor(spec.minimum, device=env.device); hi_t = torch.as_tensor(spec.maximum, device=env.device)
env.reset()
for _ in range(5): env.step(lo_t + (hi_t - lo_t) * torch.rand(N, spec.shape[0], device=env.device))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 50
e0.record()
for _ in range(K): env.step(lo_t + (hi_t - lo_t) * torch.rand(N, spec.shape[0], device=env.device))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("config 1: reach, %d envs, one B200, public API incl. goal generator hooks: %.3f ms per env step -> %.3e env-steps/s" % (N, ms, N / ms * 1e3))
