"""ctypes wrapper of the warp emulator (tests/emu/warp_emu.cpp).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "libdex_emu.so")
_CSRC = os.path.join(_DIR, "..", "..", "dexterity_b200", "csrc")
MAXCON, CON_STRIDE = 32, 8


def build(force: bool = False) -> str:
    srcs = [os.path.join(_DIR, "warp_emu.cpp")] + [
        os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", _LIB,
                               os.path.join(_DIR, "warp_emu.cpp"), "-pthread"])
    return _LIB


class EmuModel:
    def __init__(self, compiled):
        build()
        self.lib = ctypes.CDLL(_LIB)
        self.lib.emu_model_create.restype = ctypes.c_void_p
        self.lib.emu_model_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_int]
        self.lib.emu_model_destroy.argtypes = [ctypes.c_void_p]
        blob = compiled.to_blob()
        self._blob = ctypes.create_string_buffer(blob, len(blob))
        err = ctypes.create_string_buffer(256)
        self.ptr = self.lib.emu_model_create(self._blob, len(blob), err, 256)
        if not self.ptr:
            raise RuntimeError(err.value.decode())
        self.c = compiled

    def run(self, qpos, qvel, ctrl, warm, nsub=1, integrate=True, collide=True, dynamics=True,
            dtype=np.float32, debug=False, nthreads=8):
        """Returns dict with new state (+ contacts, debug arrays).  Inputs are (nenv, n) float64."""
        c = self.c
        nenv = qpos.shape[0]
        mk = lambda a, w: np.ascontiguousarray(np.pad(a, ((0, 0), (0, w - a.shape[1]))), dtype=dtype)
        Q, V, U, Wm = mk(qpos, 32), mk(qvel, 32), mk(ctrl, 32), mk(warm, 32)
        time = np.zeros(nenv, dtype)
        status = np.zeros(nenv, np.uint32)
        contact = np.zeros((nenv, MAXCON, CON_STRIDE), dtype)
        ncon = np.zeros(nenv, np.int32)
        niter = np.zeros(nenv, np.int32)
        jc = np.zeros((nenv, 3 * MAXCON, 32), dtype)
        mscr = np.zeros((nenv, 32, 32), dtype)
        xkin = np.zeros((nenv, 40 * 16), dtype)
        tree = np.zeros((nenv, 8 * 32), dtype)
        stage = np.zeros((nenv, 64, 8), dtype)
        nslot = np.zeros(nenv, np.int32)
        nv = c.nv
        dbg = dict(M=np.zeros((nenv, nv, nv), dtype), fs=np.zeros((nenv, nv), dtype), qacc=np.zeros((nenv, nv), dtype),
                   fc=np.zeros((nenv, nv), dtype), efc=np.zeros((nenv, MAXCON, 8), dtype)) if debug else {}
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
        self.lib.emu_run_steps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + \
            [ctypes.c_void_p, ctypes.c_int] * 4 + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 11 + [ctypes.c_int]
        self.lib.emu_run_steps(self.ptr, 0 if dtype == np.float32 else 1, nenv, p(Q), 32, p(V), 32, p(U), 32, p(Wm), 32,
                               p(time), p(status), p(contact), p(ncon), p(niter), nsub, int(integrate), int(collide),
                               int(dynamics), p(dbg.get("M")), p(dbg.get("fs")), p(dbg.get("qacc")), p(dbg.get("fc")),
                               p(dbg.get("efc")), p(jc), p(mscr), p(xkin), p(tree), p(stage), p(nslot), nthreads)
        out = dict(qpos=Q[:, :c.nq].astype(np.float64), qvel=V[:, :nv].astype(np.float64), warm=Wm[:, :nv].astype(np.float64),
                   time=time, status=status, contact=contact.astype(np.float64), ncon=ncon, niter=niter)
        out.update({k: v.astype(np.float64) for k, v in dbg.items()})
        return out


class EmuEnv:
    """Host emulation of dex_env_reset / dex_env_step (kernel bodies run by the fiber emulator)."""

    FIELD = dict(QPOS=0, QVEL=1, CTRL=2, QACC_WARMSTART=3, TIME=4, STATUS=5, CONTACT=6, NCON=7, NITER=8, OBS=14, REWARD=15,
                 DISCOUNT=16, NSLOT=17, GOAL=18, TASK_I=19, TASK_F=20, RNG=21, ACTION=22, EPISODE_STATS=24, RESET_MASK=25, SKIP=26, JC=27, MSCR=28, XKIN=29, TREE=30, STAGE=31, TIME64=33)

    def __init__(self, model: EmuModel, params, nenv: int, dtype=np.float32, seed: int = 1):
        self.m, self.p, self.n, self.dtype = model, params, nenv, dtype
        z = lambda *s, dt=dtype: np.zeros(s, dt)
        self.a = dict(QPOS=z(nenv, 32), QVEL=z(nenv, 32), CTRL=z(nenv, 32), QACC_WARMSTART=z(nenv, 32), TIME=z(nenv),
                      STATUS=z(nenv, dt=np.uint32), CONTACT=z(nenv, MAXCON, CON_STRIDE), NCON=z(nenv, dt=np.int32),
                      NITER=z(nenv, dt=np.int32), OBS=z(nenv, 128), REWARD=z(nenv), DISCOUNT=z(nenv), GOAL=z(nenv, 16),
                      TASK_I=z(nenv, 8, dt=np.int32), TASK_F=z(nenv, 8), RNG=z(nenv, 4, dt=np.uint32), ACTION=z(nenv, 32),
                      EPISODE_STATS=z(8, dt=np.float64), RESET_MASK=z(nenv, dt=np.uint8), SKIP=z(nenv, dt=np.uint8), JC=z(nenv, 3 * MAXCON, 32), MSCR=z(nenv, 32, 32),
                      XKIN=z(nenv, 40 * 16), TREE=z(nenv, 8 * 32), STAGE=z(nenv, 64, 8), NSLOT=z(nenv, dt=np.int32),
                      TIME64=z(nenv, 2, dt=np.float64))
        self.a["RNG"][:, 0] = seed
        self.ptrs = (ctypes.c_void_p * 40)()
        self.strides = (ctypes.c_int64 * 40)()
        for k, arr in self.a.items():
            self.ptrs[self.FIELD[k]] = arr.ctypes.data
            self.strides[self.FIELD[k]] = arr.strides[0] // arr.itemsize if arr.ndim > 1 else 1
        self.m.lib.emu_env_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]

    def _call(self, nsub, force_reset, nthreads=1):
        self.m.lib.emu_env_step(self.m.ptr, 0 if self.dtype == np.float32 else 1, self.n, self.ptrs, self.strides,
                                ctypes.byref(self.p), nsub, force_reset, nthreads)

    def reset(self): self._call(1, 1)

    def step(self, action, nsub):
        self.a["ACTION"][:, :action.shape[1]] = action
        self._call(nsub, 0, nthreads=4)


def emu_ik(model: EmuModel, params, targets: np.ndarray, dtype=np.float64, nthreads: int = 8):
    """Runs the k_ik kernel body in the emulator.  targets: [n, nsite*3]."""
    n = targets.shape[0]
    t = np.ascontiguousarray(targets, dtype=dtype)
    qout = np.zeros((n, 32), dtype)
    ok = np.zeros(n, np.uint8)
    attempts = np.zeros(n, np.int32)
    model.lib.emu_ik_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    model.lib.emu_ik_solve(model.ptr, 0 if dtype == np.float32 else 1, t.ctypes.data, n, ctypes.byref(params),
                           qout.ctypes.data, 32, ok.ctypes.data, attempts.ctypes.data, nthreads)
    return qout, ok.astype(bool), attempts
