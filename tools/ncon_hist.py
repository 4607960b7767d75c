import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import envs
N = 16384
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
H = np.zeros(25); NI = np.zeros(40); byc = {}
for it in range(40):
    env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    if it >= 10:
        nc = env.physics.ncon.cpu().numpy(); ni = env.physics.niter.cpu().numpy()
        H += np.bincount(nc, minlength=25)[:25]; NI += np.bincount(ni, minlength=40)[:40]
        for c in range(6):
            s = nc == c
            if s.any(): byc.setdefault(c, []).append(ni[s].mean())
print("ncon hist (share):", np.round(H / H.sum(), 3)[:12])
print("niter hist (share):", np.round(NI / NI.sum(), 3)[:12])
print("mean niter by ncon:", {c: round(float(np.mean(v)), 2) for c, v in byc.items()})
