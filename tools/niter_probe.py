"""Distribution of the solver's work per env (contacts, Newton iterations) in the steady state of the bench workload (GPU box)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import envs
N = 16384
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
env.physics.set_solver(int(os.environ.get("DEX_SOLVER_MODE", "0")))
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(30): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
torch.cuda.synchronize()
ncon = env.physics.ncon.cpu().numpy(); niter = env.physics.niter.cpu().numpy()
print("ncon hist", np.bincount(ncon)[:34])
for c in list(range(0, 9)) + [(9, 12), (13, 20), (21, 33)]:
    m = (ncon == c) if isinstance(c, int) else ((ncon >= c[0]) & (ncon <= c[1]))
    if m.sum(): print("ncon", c, "n", m.sum(), "niter mean %.2f p90 %d max %d" % (niter[m].mean(), np.percentile(niter[m], 90), niter[m].max()))
cost = niter * (2500 + 170 * ncon) + 3000
o = np.argsort(-cost)
print("overall niter mean %.2f; est. warp-instr per env: mean %.0f p99 %.0f max %.0f; top 10 (ncon, niter):" % (niter.mean(), cost.mean(), np.percentile(cost, 99), cost.max()),
      [(int(ncon[i]), int(niter[i])) for i in o[:10]])
print("share of total est. cost in the top 1%% / 5%% / 20%% envs: %.2f %.2f %.2f" % tuple(cost[o[:int(N * f)]].sum() / cost.sum() for f in (0.01, 0.05, 0.2)))
