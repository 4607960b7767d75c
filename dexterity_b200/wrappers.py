"""Environment wrappers (reference dexterity/manipulation/wrappers/{base,action_noise}.py)."""
from __future__ import annotations

import torch


class Wrapper:
    """Pass-through wrapper (reference wrappers/base.py:6-38)."""

    def __init__(self, environment) -> None:
        self._environment = environment

    def __getattr__(self, name):
        return getattr(self._environment, name)

    def reset(self):
        return self._environment.reset()

    def step(self, action):
        return self._environment.step(action)


class ActionNoise(Wrapper):
    """Adds clipped Gaussian noise to the action (reference wrappers/action_noise.py:9-26):
    noise ~ N(0, scale * (max - min)), action clipped back into the action spec."""

    def __init__(self, environment, scale: float = 0.01, seed: int = 0) -> None:
        super().__init__(environment)
        spec = environment.action_spec()
        dev = environment.device
        self._lo = torch.as_tensor(spec.minimum, device=dev)
        self._hi = torch.as_tensor(spec.maximum, device=dev)
        self._scale = scale
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(seed)

    def step(self, action: torch.Tensor):
        noise = torch.randn(action.shape, generator=self._gen, device=action.device, dtype=action.dtype)
        noisy = action + noise * self._scale * (self._hi - self._lo)
        return self._environment.step(torch.minimum(torch.maximum(noisy, self._lo), self._hi))
