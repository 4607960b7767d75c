import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import backend
if len(sys.argv) > 1: backend.LIB_PATH = sys.argv[1]
from dexterity_b200 import envs
env = envs.load("reorient", "state_dense", seed=12345, num_envs=16384)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(5): env.step(lo + (hi - lo) * torch.rand(16384, 20, device=env.device))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): env.step(lo + (hi - lo) * torch.rand(16384, 20, device=env.device))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(sys.argv[1:] , "env step %.3f ms -> %.3e env-steps/s niter %.2f ncon %.2f" % (ms, 16384 / ms * 1e3, env.physics.niter.float().mean().item(), env.physics.ncon.float().mean().item()))
