"""Short-horizon trajectory divergence of the FP32 CUDA path vs the FP64 oracle (north_star:
"short-horizon trajectory divergence is stated").  Same seeded states, same constant ctrl, N mj_steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dexterity_b200 import backend
from dexterity_b200.model import scenes
from tests import statelib

for scene in ("reorient_shadow", "reach_shadow", "reach_adroit"):
    c = scenes.load_compiled(scene)
    n = 256
    states = statelib.rollout_states(c, seed=777, nstates=n, every=9, episodes=32)
    model = backend.PhysicsModel(c, 0)
    print(f"## {scene}: {n} states; |dq|_inf between FP32 CUDA and FP64 oracle after k mj_steps")
    print("| k | median | p90 | max | share <= 1e-4 | contact-free at start: median |")
    print("|---|---|---|---|---|---|")
    free = None
    for k in (1, 5, 25, 100):
        ref = statelib.oracle_step(c, states, nsub=k)
        if free is None:
            free = statelib.oracle_step(c, states, nsub=1, forward_only=True)["ncon"] == 0
        b = backend.PhysicsBatch(model, n, torch.float32)
        statelib.upload(b, states)
        b.step(k); torch.cuda.synchronize()
        dq = np.abs(b.qpos.cpu().numpy().astype(np.float64) - ref["qpos"]).max(axis=1)
        print(f"| {k} | {np.median(dq):.2e} | {np.percentile(dq, 90):.2e} | {dq.max():.2e} | {(dq <= 1e-4).mean():.3f} | {np.median(dq[free]):.2e} |")
