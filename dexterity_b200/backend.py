"""ctypes binding of libdexb200.so (include/dexterity_b200.h) with zero-copy PyTorch tensors.

PyTorch is plumbing here (device memory + streams); the physics runs in the library's
hand-written sm_100a kernels.  There is no CPU fallback: importing this module without the
built library, or creating a model without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np
import torch

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libdexb200.so")

# enum DexFieldId (include/dexterity_b200.h)
FIELD = dict(QPOS=0, QVEL=1, CTRL=2, QACC_WARMSTART=3, TIME=4, STATUS=5, CONTACT=6, NCON=7, NITER=8,
             DBG_M=9, DBG_QFRC_SMOOTH=10, DBG_QACC=11, DBG_QFRC_CONSTRAINT=12, DBG_EFC=13,
             OBS=14, REWARD=15, DISCOUNT=16, NSLOT=17, GOAL=18, TASK_I=19, TASK_F=20, RNG=21, ACTION=22,
             WORK=23, EPISODE_STATS=24, RESET_MASK=25, SKIP=26, JC=27, MSCR=28, XKIN=29, TREE=30, STAGE=31, ORDER=32, TIME64=33)
STATUS_BADQACC, STATUS_CON_OVERFLOW, STATUS_CAND_OVERFLOW, STATUS_SOLVER_MAXITER, STATUS_CHOL_FAIL = 1, 2, 4, 8, 16

_lib = None


class DexError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DexError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "dexterity_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.dex_last_error.restype = ctypes.c_char_p
        L.dex_model_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.dex_model_destroy.argtypes = [ctypes.c_void_p]
        L.dex_batch_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.dex_batch_destroy.argtypes = [ctypes.c_void_p]
        L.dex_bind.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        for f in ("dex_mj_forward", "dex_mj_collide"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.dex_mj_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.dex_launch_count.restype = ctypes.c_int64
        L.dex_launch_count.argtypes = [ctypes.c_void_p]
        L.dex_capacity.argtypes = [ctypes.c_int]
        L.dex_set_chunks.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.dex_set_solver.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise DexError(f"libdexb200 error {rc}: {lib().dex_last_error().decode()}")


def capacity(which: int) -> int:
    return lib().dex_capacity(which)


class PhysicsModel:
    """Device copy of a compiled scene (replaces MjModel)."""

    def __init__(self, compiled, device: int = 0) -> None:
        self.compiled = compiled
        self.device = device
        blob = compiled.to_blob()
        self._blob = ctypes.create_string_buffer(blob, len(blob))
        self._ptr = ctypes.c_void_p()
        _check(lib().dex_model_create(self._blob, len(blob), device, ctypes.byref(self._ptr)))

    @property
    def ptr(self):
        return self._ptr

    def __del__(self):
        if getattr(self, "_ptr", None) and _lib is not None:
            _lib.dex_model_destroy(self._ptr)
            self._ptr = None


class PhysicsBatch:
    """`num_envs` independent copies of the simulation state (replaces MjData, leading env dim).

    State tensors are torch-owned and bound zero-copy; rows are padded to 32 elements so each
    warp's load/store of one env's vector is a single aligned 128-byte transaction.
    """

    ROW = 32

    def __init__(self, model: PhysicsModel, num_envs: int, dtype: torch.dtype = torch.float32,
                 debug: bool = False) -> None:
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be float32 or float64")
        c = model.compiled
        if c.nq > 48 or c.nv > self.ROW or c.nu > self.ROW:
            raise ValueError("model exceeds kernel capacities")
        self.model, self.num_envs, self.dtype = model, num_envs, dtype
        self.device = torch.device("cuda", model.device)
        self._ptr = ctypes.c_void_p()
        _check(lib().dex_batch_create(model.ptr, num_envs, 0 if dtype == torch.float32 else 1, ctypes.byref(self._ptr)))
        n, dev = num_envs, self.device
        qrow = self.ROW if c.nq <= self.ROW else 48
        maxcon, cstride = capacity(0), capacity(1)
        z = lambda *shape, dt=dtype: torch.zeros(*shape, dtype=dt, device=dev)
        self._store: Dict[str, torch.Tensor] = dict(
            QPOS=z(n, qrow), QVEL=z(n, self.ROW), CTRL=z(n, self.ROW), QACC_WARMSTART=z(n, self.ROW), TIME=z(n),
            STATUS=z(n, dt=torch.int32), CONTACT=z(n, maxcon, cstride), NCON=z(n, dt=torch.int32),
            NITER=z(n, dt=torch.int32), JC=torch.empty(n, 3 * maxcon, 32, dtype=dtype, device=dev),
            MSCR=z(n, 32, 32), XKIN=z(n, capacity(3) * 16), TREE=z(n, 8 * 32), STAGE=z(n, 64, 8),
            NSLOT=z(n, dt=torch.int32), ORDER=z(8, n, dt=torch.int32), TIME64=z(n, 2, dt=torch.float64))
        if debug:
            self._store.update(DBG_M=z(n, c.nv * c.nv), DBG_QFRC_SMOOTH=z(n, c.nv), DBG_QACC=z(n, c.nv),
                               DBG_QFRC_CONSTRAINT=z(n, c.nv), DBG_EFC=z(n, maxcon, 8))
        for name, t in self._store.items():
            self.bind(name, t)
        # global narrowphase work list: item count, fetch cursor, then (env, slot << 24 | pair) per item
        cap = max(32 * n, 4096)
        work = z(16 + 2 * cap, dt=torch.int32)
        _check(lib().dex_bind(self._ptr, FIELD["WORK"], ctypes.c_void_p(work.data_ptr()), cap))
        self._store["WORK"] = work
        self.set_chunks(int(os.environ.get("DEX_CHUNKS", "1")))
        if os.environ.get("DEX_SOLVER_MODE") or os.environ.get("DEX_SOLVER_SERIAL"):
            self.set_solver(int(os.environ.get("DEX_SOLVER_MODE", "1")))
        # public views with the model's true widths
        self.qpos = self._store["QPOS"][:, :c.nq]
        self.qvel = self._store["QVEL"][:, :c.nv]
        self.ctrl = self._store["CTRL"][:, :c.nu]
        self.qacc_warmstart = self._store["QACC_WARMSTART"][:, :c.nv]
        self.time = self._store["TIME"]
        self.time64 = self._store["TIME64"]   # [:, 0] data.time in double, [:, 1] the task's solve-start time
        self.status = self._store["STATUS"]
        self.contact = self._store["CONTACT"]
        self.ncon = self._store["NCON"]
        self.niter = self._store["NITER"]
        self.reset_to_qpos0()

    def bind(self, name: str, t: torch.Tensor) -> None:
        assert t.is_cuda and t.is_contiguous()
        stride = t.stride(0) if t.dim() > 1 else 1
        _check(lib().dex_bind(self._ptr, FIELD[name], ctypes.c_void_p(t.data_ptr()), stride))
        self._store[name] = t

    def tensor(self, name: str) -> torch.Tensor:
        return self._store[name]

    @property
    def ptr(self):
        return self._ptr

    def reset_to_qpos0(self) -> None:
        c = self.model.compiled
        self.qpos.copy_(torch.as_tensor(c.QPOS0, dtype=self.dtype, device=self.device).expand(self.num_envs, -1))
        self.qvel.zero_(); self.ctrl.zero_(); self.qacc_warmstart.zero_(); self.time.zero_(); self.time64.zero_(); self.status.zero_()

    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_chunks(self, nchunk: int) -> None:
        """Env chunks that run as independent kernel pipelines on library-owned streams (1..4)."""
        _check(lib().dex_set_chunks(self._ptr, int(nchunk)))

    def set_solver(self, mode: int) -> None:
        """0: dense warp-per-env solver for every env; 1 (default): envs without contacts take the structured
        8-lanes-per-env solver kernel; 2: also envs with <= 8 contacts of which <= 2 join two fingers; 3: sparse L^T D L form
        of the warp-per-env kernel for envs without finger-finger contacts."""
        _check(lib().dex_set_solver(self._ptr, int(mode)))

    def step(self, nsub: int = 1) -> None:
        """nsub x mj_step (collision, constraint solve, Euler integration)."""
        _check(lib().dex_mj_step(self._ptr, nsub, self._stream()))

    def forward(self) -> None:
        """mj_forward: contacts, constraint solve, no integration."""
        _check(lib().dex_mj_forward(self._ptr, self._stream()))

    def collide(self) -> None:
        _check(lib().dex_mj_collide(self._ptr, self._stream()))

    @property
    def launch_count(self) -> int:
        return int(lib().dex_launch_count(self._ptr))

    def __del__(self):
        if getattr(self, "_ptr", None) and _lib is not None:
            _lib.dex_batch_destroy(self._ptr)
            self._ptr = None
