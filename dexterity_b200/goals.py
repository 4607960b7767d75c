"""Batched goal generators mirroring the reference `GoalGenerator` interface
(dexterity/goal.py:35-73): goal_spec / initialize_episode / next_goal / relative_goal /
goal_distance / current_state / name, with a leading env dimension on every array.

* PropOrientation          ← manipulation/goals/prop_orientation.py:11-54 (torch mirror; the fused
                             reorient kernels implement the same arithmetic on device)
* FingertipCartesianPosition ← manipulation/goals/fingertip_position.py:20-139: goals are fingertip
                             positions of a joint configuration sampled around the mid-range,
                             accepted only if two settled physics steps leave the hand free of
                             self-collision.  The masked sub-rollout runs on a scratch PhysicsBatch
                             through the same CUDA kernels (dex_mj_step / dex_mj_collide).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from dexterity_b200 import backend, hands


class GoalInitializationError(RuntimeError):
    """reference dexterity/exception.py:1-2."""


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


class PropOrientation:
    name = "prop_orientation_goal_generator"

    def goal_spec(self):
        return dict(shape=(4,), dtype=torch.float64, name=self.name)

    def initialize_episode(self, physics=None, random_state=None) -> None:
        del physics, random_state

    def current_state(self, env) -> torch.Tensor:
        return env.physics.qpos[:, -4:]           # framequat of the prop geom (prop_orientation.py:31-32)

    def next_goal(self, num: int, generator: torch.Generator, device="cpu") -> torch.Tensor:
        u = torch.rand(num, 3, generator=generator, device=device)
        u1, u2, u3 = u[:, 0], 2 * math.pi * u[:, 1], 2 * math.pi * u[:, 2]    # rotations.UniformQuaternion
        return torch.stack([(1 - u1).sqrt() * u2.sin(), (1 - u1).sqrt() * u2.cos(), u1.sqrt() * u3.sin(), u1.sqrt() * u3.cos()], -1)

    def relative_goal(self, goal_state: torch.Tensor, current_state: torch.Tensor) -> torch.Tensor:
        conj = current_state * torch.tensor([1.0, -1, -1, -1], dtype=current_state.dtype, device=current_state.device)
        return quat_mul(goal_state, conj)         # tr.quat_diff_active (prop_orientation.py:40-43)

    def goal_distance(self, goal_state: torch.Tensor, current_state: torch.Tensor) -> torch.Tensor:
        w = self.relative_goal(goal_state, current_state)[..., 0].clamp(-1, 1)
        ang = 2 * torch.acos(w)
        return torch.where(ang > math.pi, 2 * math.pi - ang, ang).unsqueeze(-1)   # :45-50


class FingertipCartesianPosition:
    name = "fingertip_position_goal_generator"

    def __init__(self, hand: hands.DexterousHand, model: backend.PhysicsModel, max_rejection_samples: int = 100,
                 scale: float = 0.1) -> None:
        self._hand, self._model = hand, model
        self._max_rejection_samples, self._scale = max_rejection_samples, scale
        self._scratch: Optional[backend.PhysicsBatch] = None
        self.qpos: Optional[torch.Tensor] = None      # joint configuration behind each env's goal (:137-139)

    def goal_spec(self):
        return dict(shape=(15,), dtype=torch.float64, name=self.name)

    def _scratch_batch(self, n: int, dtype) -> backend.PhysicsBatch:
        if self._scratch is None or self._scratch.num_envs < n or self._scratch.dtype != dtype:
            self._scratch = backend.PhysicsBatch(self._model, max(n, 64), dtype)
        return self._scratch

    def current_state(self, env) -> torch.Tensor:
        return env.fingertip_positions()

    def relative_goal(self, goal_state: torch.Tensor, current_state: torch.Tensor) -> torch.Tensor:
        return goal_state - current_state             # :121-124

    def goal_distance(self, goal_state: torch.Tensor, current_state: torch.Tensor) -> torch.Tensor:
        return self.relative_goal(goal_state, current_state).reshape(*goal_state.shape[:-1], -1, 3).norm(dim=-1)  # :126-130

    def next_goal(self, env, ids: torch.Tensor, generator: torch.Generator) -> Tuple[torch.Tensor, torch.Tensor]:
        """Goals for the envs `ids` (reference :61-119).  Returns (fingertip goals [k,15], goal qpos [k,nq])."""
        k = ids.numel()
        phys = env.physics
        sc = self._scratch_batch(k, phys.dtype)
        dev, dt = phys.device, phys.dtype
        rng = torch.as_tensor(self._hand.joint_range, dtype=dt, device=dev)
        lo, hi = rng[:, 0], rng[:, 1]
        ref = 0.5 * (lo + hi)                         # reference_qpos = mid-range (:71-73)
        goals = torch.zeros(k, 3 * len(self._hand.fingertip_site_ids), dtype=dt, device=dev)
        gq = torch.zeros(k, self._hand.num_joints, dtype=dt, device=dev)
        pending = torch.arange(k, device=dev)
        for _ in range(self._max_rejection_samples):
            m = pending.numel()
            if m == 0:
                break
            src = ids[pending]
            qd = ref + self._scale * (hi - lo) * torch.randn(m, self._hand.num_joints, generator=generator, device=dev, dtype=dt)
            qd = torch.minimum(torch.maximum(qd, lo), hi)                       # np.clip (:84-86)
            sc.qpos[:m] = phys.qpos[src]; sc.qvel[:m] = phys.qvel[src]; sc.qacc_warmstart[:m] = phys.qacc_warmstart[src]
            sc.qpos[:m, :self._hand.num_joints] = qd                            # hand.set_joint_angles (:88)
            sc.ctrl[:m] = self._hand.joint_positions_to_control(qd)             # (:93-94)
            sc.status[:m] = 0
            sc.step(2)                                                          # two physics steps (:95-97)
            sc.collide()                                                        # contacts at the settled state
            ok = (sc.ncon[:m] == 0) & ((sc.status[:m] & backend.STATUS_BADQACC) == 0)   # no self collision (:99)
            acc = pending[ok]
            if acc.numel():
                tips = env.fingertip_positions(sc)[:m][ok]
                goals[acc] = tips
                gq[acc] = sc.qpos[:m, :self._hand.num_joints][ok]
            pending = pending[~ok]
        if pending.numel():
            raise GoalInitializationError(
                f"Failed to find a collision-free goal within {self._max_rejection_samples} rejection samples "
                f"for {pending.numel()} envs.")
        return goals, gq
