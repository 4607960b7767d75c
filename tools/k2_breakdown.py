"""Where K2's time goes: timing-only variants of the dense solver (variants/libdex_x*.so, built by tools/k2_breakdown.sh with
-DDEX_EXP ...) run on a FROZEN batch of bench states, so every variant sees the same workload whatever it computes.  (GPU box)"""
import ctypes, glob, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import backend, envs
N = int(os.environ.get("NENV", 16384))
torch.manual_seed(0); torch.cuda.manual_seed_all(0)
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(25): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
torch.cuda.synchronize()
ph = env.physics
blob = ph.model.compiled.to_blob()
os.environ["DEX_FREEZE"] = "1"
names = ["pre", "collide", "solve", "post", "tree", "k5", "k6", "arrow", "place", "lite"]
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for path in sorted(glob.glob("variants/libdex_x*.so")):
    L = ctypes.CDLL(os.path.abspath(path))
    L.dex_last_error.restype = ctypes.c_char_p
    L.dex_model_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.dex_batch_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.dex_bind.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    L.dex_mj_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.dex_set_solver.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dex_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dex_get_kernel_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
    buf = ctypes.create_string_buffer(blob, len(blob))
    mp, bp = ctypes.c_void_p(), ctypes.c_void_p()
    assert L.dex_model_create(buf, len(blob), 0, ctypes.byref(mp)) == 0, L.dex_last_error()
    assert L.dex_batch_create(mp, N, 0, ctypes.byref(bp)) == 0, L.dex_last_error()
    for name, t in ph._store.items():
        if name == "WORK": stride = (t.numel() - 16) // 2
        else: stride = t.stride(0) if t.dim() > 1 else 1
        assert L.dex_bind(bp, backend.FIELD[name], ctypes.c_void_p(t.data_ptr()), stride) == 0
    for mode in (0, 1):
        L.dex_set_solver(bp, mode)
        for _ in range(3): assert L.dex_mj_step(bp, 5, stream) == 0, L.dex_last_error()
        torch.cuda.synchronize()
        L.dex_set_profiling(bp, 1)
        for _ in range(4): L.dex_mj_step(bp, 5, stream)
        ms8 = (ctypes.c_double * 10)(); c8 = (ctypes.c_int64 * 10)()
        L.dex_get_kernel_times(bp, ms8, c8, 1)
        L.dex_set_profiling(bp, 0)
        print("%-28s mode %d niter %.2f | %s" % (os.path.basename(path), mode, ph.niter.float().mean().item(),
              " ".join("%s=%.3f" % (names[i], ms8[i] / max(1, c8[i])) for i in range(10) if c8[i])), flush=True)
