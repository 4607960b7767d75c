"""ctypes wrapper of the warp emulator (tests/emu/warp_emu.cpp).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "libdex_emu.so")
_CSRC = os.path.join(_DIR, "..", "..", "dexterity_b200", "csrc")
MAXCON, CON_STRIDE = 24, 8


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
        nv = c.nv
        dbg = dict(M=np.zeros((nenv, nv, nv), dtype), fs=np.zeros((nenv, nv), dtype), qacc=np.zeros((nenv, nv), dtype),
                   fc=np.zeros((nenv, nv), dtype), efc=np.zeros((nenv, MAXCON, 8), dtype)) if debug else {}
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
        self.lib.emu_run_steps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + \
            [ctypes.c_void_p, ctypes.c_int] * 4 + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 5 + [ctypes.c_int]
        self.lib.emu_run_steps(self.ptr, 0 if dtype == np.float32 else 1, nenv, p(Q), 32, p(V), 32, p(U), 32, p(Wm), 32,
                               p(time), p(status), p(contact), p(ncon), p(niter), nsub, int(integrate), int(collide),
                               int(dynamics), p(dbg.get("M")), p(dbg.get("fs")), p(dbg.get("qacc")), p(dbg.get("fc")),
                               p(dbg.get("efc")), nthreads)
        out = dict(qpos=Q[:, :c.nq].astype(np.float64), qvel=V[:, :nv].astype(np.float64), warm=Wm[:, :nv].astype(np.float64),
                   time=time, status=status, contact=contact.astype(np.float64), ncon=ncon, niter=niter)
        out.update({k: v.astype(np.float64) for k, v in dbg.items()})
        return out
