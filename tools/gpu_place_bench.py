"""Cost of the collision-checked placement rounds (GPU box): reorient env step with place_max_attempts 0 (static
clearance proof, the default for the reference's spawn box) vs PropPlacer's 20 rounds forced on the same box."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from dexterity_b200 import envs
N = int(os.environ.get("NENV", 16384))
for att in (0, 20):
    torch.manual_seed(0); torch.cuda.manual_seed_all(0)
    env = envs.BatchedGoalEnvironment("reorient", N, seed=12345, place_max_attempts=att)
    lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
    env.reset()
    for _ in range(5): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    n0 = env.physics.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    e1.record(); torch.cuda.synchronize()
    print("place_max_attempts=%d: env step %.3f ms, %d launches per step" % (att, e0.elapsed_time(e1) / 20, (env.physics.launch_count - n0) // 20))
