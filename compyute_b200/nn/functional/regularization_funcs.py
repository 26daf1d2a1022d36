"""Dropout.  API of compyute/nn/functional/regularization_funcs.py:11-52."""

from __future__ import annotations

import numpy as np

from ... import _lib, graph
from ...tensors import DeviceArray, Tensor, f32ptr, require_cuda, stream_ptr
from .functions import Function, FunctionCache, PseudoCache

__all__ = ["dropout", "DropoutFn", "set_dropout_seed"]

_seed = [0x5EED, 0]  # (seed, call counter): every call draws from a fresh counter-based stream


def set_dropout_seed(seed: int) -> None:
    _seed[0], _seed[1] = int(seed), 0


class DropoutFn(Function):
    """y = x * mask / (1-p), mask ~ Bernoulli(1-p) int8 (:15-32).  The mask comes from a device counter-based RNG,
    so values differ from NumPy's stream; the statistical contract (keep probability, scaling) is what is tested."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, p: float, training: bool) -> Tensor:
        if not training or p == 0.0:
            cache.push(False, p, None)
            return x
        require_cuda(x)
        y = DeviceArray.empty(x.shape, np.float32)
        mask = DeviceArray.empty(x.shape, np.int8)
        _seed[1] += 1
        seed = (_seed[0] * 0x9E3779B1 + _seed[1] * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF
        live = graph.replay_counter().ptr if graph.is_capturing() else None
        _lib.check(_lib.lib().cpt_dropout_fwd(f32ptr(x), y.ptr, mask.ptr, x.size, float(p), seed, live, stream_ptr()))
        cache.push(True, p, mask)
        return Tensor(y)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor) -> Tensor:
        training, p, mask = cache.pop()
        if not training:
            return dy
        dx = DeviceArray.empty(dy.shape, np.float32)
        _lib.check(_lib.lib().cpt_dropout_bwd(f32ptr(dy), mask.ptr, dx.ptr, dy.size, float(p), stream_ptr()))
        return Tensor(dx)


def dropout(x: Tensor, p: float = 0.5, training: bool = False) -> Tensor:
    return DropoutFn.forward(PseudoCache(), x, p, training)
