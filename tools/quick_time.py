"""Quick device timing of the reorient env step with the per-kernel-class breakdown (GPU box)."""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import backend
if len(sys.argv) > 1: backend.LIB_PATH = sys.argv[1]
from dexterity_b200 import envs
N = int(os.environ.get("NENV", 16384))
torch.manual_seed(0); torch.cuda.manual_seed_all(0)
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(5): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
torch.cuda.synchronize()
lib = backend.lib()
lib.dex_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.dex_get_kernel_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
lib.dex_set_profiling(env.physics.ptr, 1)
for _ in range(5): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
ms8 = (ctypes.c_double * 10)(); c8 = (ctypes.c_int64 * 10)()
lib.dex_get_kernel_times(env.physics.ptr, ms8, c8, 1)
names = ["pre", "collide", "solve", "post", "tree", "k5", "k6", "arrow", "place", "lite"]
per = " ".join("%s=%.3f" % (names[i], ms8[i] / max(1, c8[i])) for i in range(10) if c8[i])
print(sys.argv[1:], "env step %.3f ms -> %.3e env-steps/s niter %.2f ncon %.2f | per launch ms: %s" % (
    ms, N / ms * 1e3, env.physics.niter.float().mean().item(), env.physics.ncon.float().mean().item(), per))
