"""Solver workload of the bench configuration: K1c class counts, Newton iterations by contact count, and the solver
kernels timed one after the other (DEX_SOLVER_SERIAL) in the three solver modes.  (GPU box)"""
import ctypes, os, sys
os.environ["DEX_SOLVER_SERIAL"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import backend, envs
N = int(os.environ.get("NENV", 16384))
torch.manual_seed(0); torch.cuda.manual_seed_all(0)
lib = backend.lib()
lib.dex_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.dex_get_kernel_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
names = ["pre", "collide", "solve", "post", "tree", "k5", "k6", "arrow", "place", "lite"]
for mode in (0, 1, 2):
    env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
    env.physics.set_solver(mode)
    lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
    env.reset()
    for _ in range(25): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    torch.cuda.synchronize()
    lib.dex_set_profiling(env.physics.ptr, 1)
    for _ in range(5): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    ms8 = (ctypes.c_double * 10)(); c8 = (ctypes.c_int64 * 10)()
    lib.dex_get_kernel_times(env.physics.ptr, ms8, c8, 1)
    lib.dex_set_profiling(env.physics.ptr, 0)
    print("mode", mode, "serial per launch ms:", " ".join("%s=%.3f" % (names[i], ms8[i] / max(1, c8[i])) for i in range(10) if c8[i]))
    if mode == 2:
        w = env.physics.tensor("WORK").cpu().numpy()
        print("class counts (0-3 dense by ncon>=12/9/4/else, 4 cross, 5 ncon>=3, 6 ncon>=1, 7 free):", w[2:10].tolist(), "narrow items", int(w[0]))
        nc = env.physics.ncon.cpu().numpy(); ni = env.physics.niter.cpu().numpy()
        print("ncon hist", np.bincount(nc, minlength=12)[:20].tolist())
        print("niter hist", np.bincount(ni, minlength=12)[:20].tolist())
        for k in range(0, 10):
            sel = nc == k
            if sel.any(): print("  ncon %d: %5d envs, niter mean %.2f max %d" % (k, sel.sum(), ni[sel].mean(), ni[sel].max()))
    del env
