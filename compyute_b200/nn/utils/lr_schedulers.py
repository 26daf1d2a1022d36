"""Learning-rate schedulers (API of compyute/nn/utils/lr_schedulers.py:27-217).

Host-side policy only: a scheduler rewrites ``optimizer.lr`` (a Python float) between steps as a function of the optimizer's
step counter ``t``.  On the device nothing changes shape: the fused update kernels take ``lr`` as a launch argument, and a
CUDA-graph-captured step reads it from the live-scalar slot that ``Optimizer.step`` refreshes before every replay
(``cpt_set_live_scalars``) — so schedulers keep working under graph replay (tests/test_gpu_models.py).
Every ``step()`` first appends the current rate to ``cache["lr_history"]`` like the reference.
"""

from __future__ import annotations

import math
from typing import Any, Callable

__all__ = ["LrScheduler", "StepLrScheduler", "MultistepLrScheduler", "ExponentialLrScheduler", "CosineLrScheduler",
           "AdaptiveLrScheduler"]


class LrScheduler:
    """Base: subclasses give ``_rule(t, lr, **metrics) -> new lr`` (lr_schedulers.py:27-45)."""

    def __init__(self, optimizer) -> None:
        self.optimizer = optimizer
        self.cache: dict[str, list[float]] = {"lr_history": []}

    def _rule(self, t: int, lr: float, **metrics: Any) -> float:
        raise NotImplementedError

    def step(self, **kwargs: Any) -> None:
        self.cache["lr_history"].append(self.optimizer.lr)
        self.optimizer.lr = self._rule(self.optimizer.t, self.optimizer.lr, **kwargs)


def _scaled_when(cond: Callable[[int], bool]):
    """lr *= lr_decay at the steps where ``cond(t - 1)`` holds (``t`` starts at 1, optimizers.py:29)."""

    def rule(self, t: int, lr: float, **_: Any) -> float:
        return lr * self.lr_decay if cond(self, t - 1) else lr

    return rule


class StepLrScheduler(LrScheduler):
    """One decay after ``t_decay`` steps (lr_schedulers.py:48-71)."""

    def __init__(self, optimizer, t_decay: int, lr_decay: float = 0.1) -> None:
        super().__init__(optimizer)
        self.t_decay, self.lr_decay = t_decay, lr_decay

    _rule = _scaled_when(lambda self, done: done == self.t_decay)


class MultistepLrScheduler(LrScheduler):
    """A decay every ``t_decay_step`` steps (lr_schedulers.py:74-97)."""

    def __init__(self, optimizer, t_decay_step: int, lr_decay: float = 0.1) -> None:
        super().__init__(optimizer)
        self.t_decay_step, self.lr_decay = t_decay_step, lr_decay

    _rule = _scaled_when(lambda self, done: done % self.t_decay_step == 0)


class ExponentialLrScheduler(LrScheduler):
    """A decay on each of the first ``decay_steps`` steps (lr_schedulers.py:100-123)."""

    def __init__(self, optimizer, decay_steps: int, lr_decay: float = 0.1) -> None:
        super().__init__(optimizer)
        self.decay_steps, self.lr_decay = decay_steps, lr_decay

    _rule = _scaled_when(lambda self, done: done <= self.decay_steps)


class CosineLrScheduler(LrScheduler):
    """Linear warm-up to the optimizer's initial rate, half-cosine down to ``target_lr``, then flat (lr_schedulers.py:126-170)."""

    def __init__(self, optimizer, target_lr: float, warmup_steps: int, decay_steps: int) -> None:
        super().__init__(optimizer)
        self.target_lr, self.warmup_steps, self.decay_steps = target_lr, warmup_steps, decay_steps
        self._max_lr = optimizer.lr

    def _rule(self, t: int, lr: float, **_: Any) -> float:
        if t <= self.warmup_steps:
            return self._max_lr / self.warmup_steps * t
        if t > self.warmup_steps + self.decay_steps:
            return self.target_lr
        phase = (t - self.warmup_steps) / self.decay_steps
        return self.target_lr + 0.5 * (1.0 + math.cos(math.pi * phase)) * (self._max_lr - self.target_lr)


class AdaptiveLrScheduler(LrScheduler):
    """Scales the rate by the trend of one metric over the last ``patience`` steps (lr_schedulers.py:173-217): a falling
    metric multiplies by ``lr_upscale_factor``, anything else by ``lr_downscale_factor``."""

    def __init__(self, optimizer, patience: int = 10, lr_downscale_factor: float = 0.1, lr_upscale_factor: float = 2.0) -> None:
        super().__init__(optimizer)
        self.patience, self.lr_downscale_factor, self.lr_upscale_factor = patience, lr_downscale_factor, lr_upscale_factor

    def step(self, **kwargs: Any) -> None:
        if len(kwargs) != 1:
            raise ValueError("Exactly one metric value must be passed as kwarg.")
        super().step(**kwargs)

    def _rule(self, t: int, lr: float, **metrics: Any) -> float:
        hist = self.cache.setdefault("target_history", [])
        hist.append(next(iter(metrics.values())))
        if t <= self.patience:
            return lr
        window = hist[-self.patience - 1:]
        trend = sum(b - a for a, b in zip(window, window[1:]))
        return lr * (self.lr_upscale_factor if trend < 0 else self.lr_downscale_factor)
