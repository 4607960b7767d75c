import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import envs
env = envs.load("reorient", "state_dense", seed=12345, num_envs=16384)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(30): env.step(lo + (hi - lo) * torch.rand(16384, 20, device=env.device))
torch.cuda.synchronize()
ncon = env.physics.ncon.cpu().numpy(); niter = env.physics.niter.cpu().numpy()
print("ncon hist", np.bincount(ncon)[:12])
for c in range(0, 8):
    m = ncon == c
    if m.sum(): print("ncon", c, "n", m.sum(), "niter mean %.2f p90 %d max %d" % (niter[m].mean(), np.percentile(niter[m], 90), niter[m].max()), "hist", np.bincount(niter[m])[:10])
print("overall niter mean", niter.mean())
