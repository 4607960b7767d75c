"""ctypes wrapper of the FP64 CPU oracle (oracle/mjstep_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
PARITY UNPINNED (see the .cpp header and DESIGN.md): MuJoCo itself is not available.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "mjstep_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_model_create.restype = ctypes.c_void_p
        L.orc_model_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        L.orc_model_destroy.argtypes = [ctypes.c_void_p]
        L.orc_data_create.restype = ctypes.c_void_p
        L.orc_data_create.argtypes = [ctypes.c_void_p]
        L.orc_data_destroy.argtypes = [ctypes.c_void_p]
        L.orc_set_state.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, ctypes.c_double]
        L.orc_get_state.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp]
        for f in ("orc_forward", "orc_step", "orc_step1", "orc_step2", "orc_kinematics"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.orc_get.restype = ctypes.c_int64
        L.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _dp, ctypes.c_int64]
        L.orc_nefc.argtypes = [ctypes.c_void_p]
        L.orc_solver_niter.argtypes = [ctypes.c_void_p]
        L.orc_efc_types.argtypes = [ctypes.c_void_p, _ip, _ip, ctypes.c_int]
        L.orc_contacts.argtypes = [ctypes.c_void_p, _ip, _dp, ctypes.c_int]
        L.orc_jac.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp]
        L.orc_object_velocity.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.orc_mpr.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp]
        L.orc_support.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.orc_step_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, _dp, _dp, _dp, _dp, _ip,
                                     ctypes.c_int]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


class OracleModel:
    def __init__(self, compiled) -> None:
        self.compiled = compiled
        blob = compiled.to_blob()
        self._blob = ctypes.create_string_buffer(blob, len(blob))
        self.ptr = lib().orc_model_create(self._blob, len(blob))
        if not self.ptr:
            raise RuntimeError("oracle rejected the model blob")
        self.nq, self.nv, self.nu = compiled.nq, compiled.nv, compiled.nu

    def __del__(self):
        if getattr(self, "ptr", None) and lib is not None:   # `lib` is already gone at interpreter shutdown
            lib().orc_model_destroy(self.ptr)
            self.ptr = None

    def step_batch(self, qpos, qvel, ctrl, warm, nsub: int, nthreads: int = 1):
        """In-place batched env step: nsub mj_steps + mj_step1 refresh per env."""
        n = qpos.shape[0]
        for a in (qpos, qvel, ctrl, warm):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        ncon = np.zeros(n, dtype=np.int32)
        lib().orc_step_batch(self.ptr, n, nsub, _p(qpos), _p(qvel), _p(ctrl), _p(warm),
                             ncon.ctypes.data_as(_ip), nthreads)
        return ncon


class OracleData:
    """One environment (the analogue of an MjData)."""

    def __init__(self, model: OracleModel) -> None:
        self.model = model
        self.ptr = lib().orc_data_create(model.ptr)

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_data_destroy(self.ptr)
            self.ptr = None

    def set_state(self, qpos=None, qvel=None, ctrl=None, qacc_warmstart=None, time: float = 0.0) -> None:
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        qpos, qvel, ctrl, w = c(qpos), c(qvel), c(ctrl), c(qacc_warmstart)
        lib().orc_set_state(self.ptr, _p(qpos), _p(qvel), _p(ctrl), _p(w), time)

    def get_state(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, float]:
        m = self.model
        qpos, qvel, warm = np.zeros(m.nq), np.zeros(m.nv), np.zeros(m.nv)
        t = ctypes.c_double()
        lib().orc_get_state(self.ptr, _p(qpos), _p(qvel), _p(warm), ctypes.byref(t))
        return qpos, qvel, warm, t.value

    def forward(self): lib().orc_forward(self.ptr)
    def step(self): lib().orc_step(self.ptr)
    def step1(self): lib().orc_step1(self.ptr)
    def step2(self): lib().orc_step2(self.ptr)
    def kinematics(self): lib().orc_kinematics(self.ptr)

    def get(self, name: str) -> np.ndarray:
        n = lib().orc_get(self.ptr, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        lib().orc_get(self.ptr, name.encode(), _p(out), n)
        return out

    @property
    def nefc(self) -> int:
        return lib().orc_nefc(self.ptr)

    @property
    def solver_niter(self) -> int:
        return lib().orc_solver_niter(self.ptr)

    def efc_types(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self.nefc
        t, i = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lib().orc_efc_types(self.ptr, t.ctypes.data_as(_ip), i.ctypes.data_as(_ip), n)
        return t, i

    def contacts(self, cap: int = 256) -> Dict[str, np.ndarray]:
        g = np.zeros((cap, 2), np.int32)
        v = np.zeros((cap, 13))
        n = lib().orc_contacts(self.ptr, g.ctypes.data_as(_ip), _p(v), cap)
        n = min(n, cap)
        return dict(geom=g[:n], dist=v[:n, 0], pos=v[:n, 1:4], frame=v[:n, 4:13].reshape(n, 3, 3))

    def jac(self, body: int, point: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        nv = self.model.nv
        jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
        pt = np.ascontiguousarray(point, dtype=np.float64)
        lib().orc_jac(self.ptr, body, _p(pt), _p(jp), _p(jr))
        return jp, jr

    def object_velocity(self, body: int, point: np.ndarray) -> np.ndarray:
        out = np.zeros(6)
        pt = np.ascontiguousarray(point, dtype=np.float64)
        lib().orc_object_velocity(self.ptr, body, _p(pt), _p(out))
        return out

    def mpr(self, g1: int, g2: int):
        depth = ctypes.c_double()
        d, p = np.zeros(3), np.zeros(3)
        r = lib().orc_mpr(self.ptr, g1, g2, ctypes.byref(depth), _p(d), _p(p))
        return r, depth.value, d, p

    def support(self, g: int, direction: np.ndarray) -> np.ndarray:
        out = np.zeros(3)
        dd = np.ascontiguousarray(direction, dtype=np.float64)
        lib().orc_support(self.ptr, g, _p(dd), _p(out))
        return out
