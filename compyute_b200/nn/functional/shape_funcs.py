"""Flatten.  API of compyute/nn/functional/shape_funcs.py:8-51 — a zero-copy view on C-contiguous NCHW data."""

from __future__ import annotations

from ...tensors import Tensor
from .functions import Function, FunctionCache, PseudoCache

__all__ = ["flatten", "FlattenFn"]


class FlattenFn(Function):
    @staticmethod
    def forward(cache: FunctionCache, x: Tensor) -> Tensor:
        cache.push(x.shape)
        return x.view((x.shape[0], -1))

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor) -> Tensor:
        (shape,) = cache.pop()
        return dy.view(shape)


def flatten(x: Tensor) -> Tensor:
    return FlattenFn.forward(PseudoCache(), x)
