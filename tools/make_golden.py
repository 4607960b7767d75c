"""Writes tests/golden/reorient_oracle_steps.npz: seeded states and the FP64 oracle's one-step
results.  These are NOT MuJoCo outputs (MuJoCo cannot be installed here); they pin the oracle
against accidental change.  When a machine with `mujoco` + `dm_control` is available, the same
state arrays can be replayed through the reference (scripts/export_task.py exports the MJCF) and
the `next_*` arrays replaced by true mj_step outputs — the parity tests consume either."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dexterity_b200.model import scenes
from tests import statelib

c = scenes.load_compiled("reorient_shadow")
states = statelib.rollout_states(c, seed=4242, nstates=48, every=9, episodes=6)
ref = statelib.oracle_step(c, states, nsub=1)
out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reorient_oracle_steps.npz")
np.savez_compressed(out, **states, next_qpos=ref["qpos"], next_qvel=ref["qvel"], next_warm=ref["warm"],
                    ncon=ref["ncon"], qacc=ref["qacc"])
print("wrote", out, os.path.getsize(out), "bytes; ncon hist", np.bincount(ref["ncon"]))
