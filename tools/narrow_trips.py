"""Cycles and active lanes per trip of K1b's per-warp loop (variants/libdex_nprof.so, -DDEX_NARROW_PROF).  (GPU box)"""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import backend
backend.LIB_PATH = "variants/libdex_nprof.so"
from dexterity_b200 import envs
N = 16384
torch.manual_seed(0)
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(25): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
lib = backend.lib()
out = (ctypes.c_ulonglong * 384)()
lib.dex_debug_trips(out, 1)
for _ in range(4): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
lib.dex_debug_trips(out, 0)
a = np.array(out[:], dtype=np.float64).reshape(3, 128)
launches = 20
cyc, warps, lanes = a
tot = cyc.sum()
print("warps launched per kernel", warps[0] / launches, "total warp-cycles per launch %.3e" % (tot / launches))
acc = 0
for t in range(128):
    if warps[t] == 0: continue
    acc += cyc[t]
    if t < 12 or t % 8 == 0 or t == 127:
        print("trip %3d: warps alive %6.0f  lanes/warp %5.2f  cycles/warp-trip %7.0f  share of warp-cycles %5.1f%%  cumulative %5.1f%%" % (
            t, warps[t] / launches, lanes[t] / warps[t], cyc[t] / warps[t], 100 * cyc[t] / tot, 100 * acc / tot))
