"""GPU probe (development aid): parity statistics vs the oracle and raw kernel timings."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dexterity_b200 import backend
from dexterity_b200.model import scenes
from tests import statelib

compiled = scenes.load_compiled("reorient_shadow")
print("cpu cores", os.cpu_count(), "gpu", torch.cuda.get_device_name(0))
model = backend.PhysicsModel(compiled, 0)
N = int(os.environ.get("NSTATES", 256))
t0 = time.time()
states = statelib.rollout_states(compiled, seed=123, nstates=N, every=9, episodes=max(1, N // 16))
ref = statelib.oracle_step(compiled, states, nsub=1)
print("oracle states+ref %.1fs" % (time.time() - t0), "ncon hist", np.bincount(ref["ncon"]))
for dtype in (torch.float64, torch.float32):
    b = backend.PhysicsBatch(model, N, dtype, debug=True)
    statelib.upload(b, states)
    b.step(1); torch.cuda.synchronize()
    q = b.qpos.cpu().numpy().astype(np.float64); v = b.qvel.cpu().numpy().astype(np.float64)
    ncon = b.ncon.cpu().numpy(); st = b.status.cpu().numpy()
    eq = np.abs(q - ref["qpos"]).max(axis=1) / np.maximum(1.0, np.abs(ref["qpos"]).max(axis=1))
    ev = np.abs(v - ref["qvel"]).max(axis=1) / np.maximum(1.0, np.abs(ref["qvel"]).max(axis=1))
    same = ncon == ref["ncon"]
    print(str(dtype), "ncon equal %d/%d" % (same.sum(), N), "status nonzero", int((st != 0).sum()), np.unique(st),
          "relerr qpos med %.2e p90 %.2e max %.2e" % (np.median(eq), np.percentile(eq, 90), eq.max()),
          "qvel med %.2e p90 %.2e max %.2e" % (np.median(ev), np.percentile(ev, 90), ev.max()),
          "frac qvel<=1e-4: %.3f" % (ev <= 1e-4).mean(), "niter mean %.2f max %d" % (b.niter.float().mean().item(), b.niter.max().item()))
    del b

# ---- timing
for nenv in (4096, 16384):
    b = backend.PhysicsBatch(model, nenv, torch.float32)
    idx = np.arange(nenv) % N
    statelib.upload(b, {k: s[idx] for k, s in states.items()})
    lo = torch.as_tensor(compiled.ACT_CTRLRANGE[:, 0], dtype=torch.float32, device="cuda")
    hi = torch.as_tensor(compiled.ACT_CTRLRANGE[:, 1], dtype=torch.float32, device="cuda")
    for w in range(3):
        b.ctrl.copy_(lo + (hi - lo) * torch.rand(nenv, compiled.nu, device="cuda")); b.step(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for k in range(K):
        b.ctrl.copy_(lo + (hi - lo) * torch.rand(nenv, compiled.nu, device="cuda")); b.step(5)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("nenv %d: %.3f ms per env-step(5 substeps) -> %.3e env-steps/s, %.3e mj_steps/s; ncon mean %.2f niter mean %.2f status %s" % (
        nenv, ms, nenv / ms * 1e3, 5 * nenv / ms * 1e3, b.ncon.float().mean().item(), b.niter.float().mean().item(),
        np.unique(b.status.cpu().numpy())))
    # split timing: collide only vs full
    e0.record()
    for k in range(K * 5): b.collide()
    e1.record(); torch.cuda.synchronize()
    print("   collide-only: %.3f ms per launch" % (e0.elapsed_time(e1) / (K * 5)))
    e0.record()
    for k in range(K * 5): b.forward()
    e1.record(); torch.cuda.synchronize()
    print("   forward (collide+dynamics, no integrate): %.3f ms per launch" % (e0.elapsed_time(e1) / (K * 5)))
    del b
