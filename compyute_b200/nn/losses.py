"""Loss / metric classes.  ``CrossEntropyLoss`` (compyute/nn/losses.py:129-146), ``Accuracy`` (metrics.py:33)."""

from __future__ import annotations

from ..tensors import Tensor
from .functional.functions import FunctionCache
from .functional.loss_funcs import CrossEntropyLossFn, accuracy_score

__all__ = ["Loss", "CrossEntropyLoss", "Accuracy"]


class Loss:
    """Loss base: owns a FunctionCache; ``__call__`` = forward (losses.py:23-106)."""

    def __init__(self) -> None:
        self.fcache = FunctionCache()
        self.label = self.__class__.__name__

    def __call__(self, logits: Tensor, targets: Tensor) -> Tensor:
        return self.forward(logits, targets)


class CrossEntropyLoss(Loss):
    def forward(self, logits: Tensor, targets: Tensor, eta: float = 1e-8) -> Tensor:
        return CrossEntropyLossFn.forward(self.fcache, logits, targets, eta)

    def backward(self) -> Tensor:
        return CrossEntropyLossFn.backward(self.fcache)


class Accuracy:
    def __call__(self, logits: Tensor, targets: Tensor) -> float:
        return accuracy_score(logits, targets)
