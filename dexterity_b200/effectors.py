"""Batched effectors (reference dexterity/effector.py:8-34, effectors/mujoco_actuation.py:11-67,
effectors/hand_effector.py:10-37, effectors/wrappers/{previous_action,smooth_action}.py).

`set_control(physics, command)` keeps the reference signature; `physics` is a PhysicsBatch and
`command` a [num_envs, n] tensor.  The registered tasks run the plain effector inside the fused
pre-step kernel; these classes serve custom tasks and the optional wrappers.
"""
from __future__ import annotations

import abc
from typing import Optional, Sequence

import numpy as np
import torch

from dexterity_b200.envs import BoundedArraySpec


class Effector(abc.ABC):
    @abc.abstractmethod
    def after_compile(self, model) -> None: ...
    @abc.abstractmethod
    def initialize_episode(self, physics, random_state=None) -> None: ...
    @abc.abstractmethod
    def action_spec(self, physics) -> BoundedArraySpec: ...
    @abc.abstractmethod
    def set_control(self, physics, command: torch.Tensor) -> None: ...
    @property
    @abc.abstractmethod
    def prefix(self) -> str: ...


def create_action_spec(compiled, actuator_ids: Sequence[int], prefix: str = "") -> BoundedArraySpec:
    """reference effectors/mujoco_actuation.py:48-67."""
    n = len(actuator_ids)
    lo = np.full(n, -np.inf, np.float32)
    hi = np.full(n, np.inf, np.float32)
    ids = np.asarray(actuator_ids)
    lim = compiled.ACT_CTRLLIMITED[ids].astype(bool)
    lo[lim], hi[lim] = compiled.ACT_CTRLRANGE[ids][lim, 0], compiled.ACT_CTRLRANGE[ids][lim, 1]
    return BoundedArraySpec((n,), np.dtype(np.float32), lo, hi, "\t".join(f"{prefix}{i}" for i in range(n)))


class MujocoEffector(Effector):
    def __init__(self, compiled, actuator_ids: Sequence[int], prefix: str = "") -> None:
        self._compiled, self._ids, self._prefix = compiled, list(actuator_ids), prefix
        self._spec: Optional[BoundedArraySpec] = None

    def after_compile(self, model) -> None: del model
    def initialize_episode(self, physics, random_state=None) -> None: del physics, random_state

    def action_spec(self, physics=None) -> BoundedArraySpec:
        if self._spec is None:
            self._spec = create_action_spec(self._compiled, self._ids, self._prefix)
        return self._spec

    def set_control(self, physics, command: torch.Tensor) -> None:
        # MuJoCo clips out-of-range controls itself; the stored ctrl stays unclipped (mujoco_actuation.py:30-33)
        physics.ctrl[:, self._ids] = command.to(physics.ctrl.dtype)

    @property
    def prefix(self) -> str: return self._prefix


class HandEffector(MujocoEffector):
    """reference effectors/hand_effector.py:10-37 (prefix '<hand>_joint')."""

    def __init__(self, hand, hand_name: str) -> None:
        super().__init__(hand.compiled, hand.actuator_ids, prefix=f"{hand_name}_joint")


class _Wrapper(Effector):
    def __init__(self, effector: Effector) -> None: self._effector = effector
    def after_compile(self, model) -> None: self._effector.after_compile(model)
    def action_spec(self, physics=None): return self._effector.action_spec(physics)
    @property
    def prefix(self) -> str: return self._effector.prefix


class PreviousAction(_Wrapper):
    """Remembers the last command (a copy); reference effectors/wrappers/previous_action.py:10-34."""

    def __init__(self, effector: Effector) -> None:
        super().__init__(effector)
        self._previous_action: Optional[torch.Tensor] = None

    def initialize_episode(self, physics, random_state=None) -> None:
        self._effector.initialize_episode(physics, random_state)
        spec = self.action_spec(physics)
        self._previous_action = torch.zeros(physics.num_envs, spec.shape[0], dtype=physics.dtype, device=physics.device)

    def set_control(self, physics, command: torch.Tensor) -> None:
        self._previous_action = command.clone()
        self._effector.set_control(physics, command)

    @property
    def previous_action(self) -> torch.Tensor: return self._previous_action


class SmoothAction(_Wrapper):
    """Exponential moving average v <- alpha v + (1 - alpha) x, first call passes x through;
    reference effectors/wrappers/smooth_action.py:10-63."""

    def __init__(self, effector: Effector, alpha: float = 0.5) -> None:
        if not 0.0 <= alpha <= 1.0:
            raise ValueError("`alpha` must be in [0, 1].")
        super().__init__(effector)
        self._alpha = alpha
        self._value: Optional[torch.Tensor] = None

    def initialize_episode(self, physics, random_state=None) -> None:
        self._effector.initialize_episode(physics, random_state)
        self._value = None

    def set_control(self, physics, command: torch.Tensor) -> None:
        self._value = command.clone() if self._value is None else self._alpha * self._value + (1 - self._alpha) * command
        self._effector.set_control(physics, self._value)
