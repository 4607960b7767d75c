"""Batched mirror of the reference `IKSolver` (dexterity/inverse_kinematics/ik_solver.py:34-153).

`IKSolver(hand).solve(target_positions, linear_tol, max_steps, early_stop, num_attempts,
stop_on_first_successful_attempt)` keeps the reference signature: a single (num_fingers, 3) target
returns a joint configuration or None; a batch [n, num_fingers, 3] (numpy or CUDA tensor) returns
(qpos [n, nq], ok [n]).  The solve runs in the library's `k_ik` kernel, one warp per target set.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple, Union

import numpy as np
import torch

from dexterity_b200 import backend, hands
from dexterity_b200.model import scenes


class IkParams(ctypes.Structure):
    """struct DexIkParams (include/dexterity_b200.h)."""
    _fields_ = [("nsite", ctypes.c_int), ("site_id", ctypes.c_int * 5), ("max_steps", ctypes.c_int),
                ("num_attempts", ctypes.c_int), ("early_stop", ctypes.c_int), ("stop_on_first", ctypes.c_int),
                ("linear_tol", ctypes.c_double), ("gain", ctypes.c_double), ("regularization", ctypes.c_double),
                ("progress_threshold", ctypes.c_double), ("seed", ctypes.c_ulonglong)]


# reference ik_solver.py:20-31
_LINEAR_VELOCITY_GAIN = 0.95
_REGULARIZATION_WEIGHT = 1e-5
_PROGRESS_THRESHOLD = 20.0


class IKSolver:
    """Inverse kinematics solver for a dexterous hand (private kinematic model, like the reference)."""

    def __init__(self, hand: Optional[hands.DexterousHand] = None, device: int = 0, seed: int = 0,
                 hand_name: str = "shadow_hand_e") -> None:
        """`hand_name`: "shadow_hand_e" or "adroit_hand" (reference models/hands/{shadow_hand_e,adroit_hand}.py)."""
        if hand_name not in scenes.HANDS:
            raise ValueError(f"Unknown hand: {hand_name}")
        asset, prefix, tips = scenes.HANDS[hand_name]
        compiled = scenes.load_compiled(asset)
        self.hand = hand or hands.DexterousHand(compiled, prefix, tips)
        self.compiled = compiled
        self.model = backend.PhysicsModel(compiled, device)
        self.device = torch.device("cuda", device)
        self._sites = [compiled.site_names.index(prefix + t) for t in tips]
        self._seed = seed
        self._calls = 0
        lib = backend.lib()
        lib.dex_ik_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(IkParams),
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    @property
    def num_sites(self) -> int:
        return len(self._sites)

    def solve(self, target_positions: Union[np.ndarray, torch.Tensor], linear_tol: float = 1e-3, max_steps: int = 100,
              early_stop: bool = False, num_attempts: int = 30, stop_on_first_successful_attempt: bool = False,
              dtype: torch.dtype = torch.float32):
        single = not torch.is_tensor(target_positions) and np.asarray(target_positions).size == 3 * self.num_sites \
            and np.asarray(target_positions).ndim <= 2
        t = torch.as_tensor(np.asarray(target_positions) if not torch.is_tensor(target_positions) else target_positions)
        if t.numel() % (3 * self.num_sites) != 0 or (t.dim() >= 2 and t.shape[-1] != 3) or \
                (t.dim() >= 2 and t.shape[-2] != self.num_sites and t.shape[-1] * t.shape[-2] != 3 * self.num_sites):
            raise ValueError("The number of target positions must be equal to the number of end-effector sites.")
        t = t.reshape(-1, 3 * self.num_sites).to(device=self.device, dtype=dtype).contiguous()
        n = t.shape[0]
        p = IkParams()
        p.nsite = self.num_sites
        for i, s in enumerate(self._sites):
            p.site_id[i] = s
        p.max_steps, p.num_attempts = int(max_steps), int(num_attempts)
        p.early_stop, p.stop_on_first = int(early_stop), int(stop_on_first_successful_attempt)
        p.linear_tol, p.gain = float(linear_tol), _LINEAR_VELOCITY_GAIN
        p.regularization, p.progress_threshold = _REGULARIZATION_WEIGHT, _PROGRESS_THRESHOLD
        self._calls += 1
        p.seed = (self._seed << 20) + self._calls
        qout = torch.zeros(n, 32, dtype=dtype, device=self.device)
        ok = torch.zeros(n, dtype=torch.uint8, device=self.device)
        attempts = torch.zeros(n, dtype=torch.int32, device=self.device)
        backend._check(backend.lib().dex_ik_solve(
            self.model.ptr, 0 if dtype == torch.float32 else 1, ctypes.c_void_p(t.data_ptr()), n, ctypes.byref(p),
            ctypes.c_void_p(qout.data_ptr()), 32, ctypes.c_void_p(ok.data_ptr()), ctypes.c_void_p(attempts.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        self.last_attempts = attempts
        q = qout[:, :self.compiled.nq]
        if single:
            return q[0].cpu().numpy().astype(np.float64) if bool(ok[0].item()) else None
        return q, ok.bool()
