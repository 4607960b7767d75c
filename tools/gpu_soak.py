"""Soak: 600 control steps (3000 mj_steps) of reorient at 16384 envs with random actions and auto-reset; reports
status bits, non-finite state and episode statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dexterity_b200 import envs
N = 16384
torch.manual_seed(1)
env = envs.load("reorient", "state_dense", seed=777, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
bits = torch.zeros(N, dtype=torch.int32, device=env.device)
nonfinite = 0
for it in range(600):
    ts = env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    bits |= env.physics.status
    if it % 50 == 49:
        nonfinite += int((~torch.isfinite(env.physics.qpos).all(dim=1)).sum().item()) + int((~torch.isfinite(ts.reward)).sum().item())
names = {1: "BADQACC", 2: "CON_OVERFLOW", 4: "CAND_OVERFLOW", 8: "SOLVER_MAXITER", 16: "CHOL_FAIL", 32: "PLACEMENT"}
print("envs with status bit ever set over 600 steps:", {n: int(((bits & b) != 0).sum().item()) for b, n in names.items()})
print("non-finite samples:", nonfinite)
print("episode stats:", env.reduce_episode_stats())
