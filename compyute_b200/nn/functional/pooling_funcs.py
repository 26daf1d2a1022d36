"""MaxPooling2D / AvgPooling2D.  API of compyute/nn/functional/pooling_funcs.py:67-143."""

from __future__ import annotations

import numpy as np

from ... import _lib
from ...tensors import DeviceArray, ShapeError, Tensor, f32ptr, require_cuda, stream_ptr
from .activation_funcs import FUSED_INTO_PRODUCER
from .functions import Function, FunctionCache, PseudoCache

__all__ = ["maxpooling2d", "avgpooling2d", "MaxPooling2DFn", "AvgPooling2DFn"]


def _check4d(x: Tensor) -> None:
    if x.ndim != 4:
        raise ShapeError(f"Expected input to be 4D, got {x.ndim}D.")


class MaxPooling2DFn(Function):
    """Non-overlapping k x k max (stride = k, floor); backward = equality mask, ties all receive dy (:71-82)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, kernel_size: int) -> Tensor:
        _check4d(x)
        require_cuda(x)
        B, C, H, W = x.shape
        k = int(kernel_size)
        y = DeviceArray.empty((B, C, H // k, W // k), np.float32)
        _lib.check(_lib.lib().cpt_maxpool2d_fwd(f32ptr(x), y.ptr, B, C, H, W, k, stream_ptr()))
        yt = Tensor(y)
        cache.push(x, k, yt)
        return yt

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor) -> Tensor:
        x, k, y = cache.pop()
        if x is FUSED_INTO_PRODUCER:  # evaluated inside the preceding BatchNorm (BatchNorm2D -> ReLU -> MaxPooling2D(2) peephole)
            return dy
        require_cuda(dy)
        B, C, H, W = x.shape
        dx = DeviceArray.empty(x.shape, np.float32)
        _lib.check(_lib.lib().cpt_maxpool2d_bwd(f32ptr(x), f32ptr(y), f32ptr(dy), dx.ptr, B, C, H, W, k, stream_ptr()))
        return Tensor(dx)


class AvgPooling2DFn(Function):
    """Non-overlapping k x k mean (:107-121)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, kernel_size: int) -> Tensor:
        _check4d(x)
        require_cuda(x)
        B, C, H, W = x.shape
        k = int(kernel_size)
        y = DeviceArray.empty((B, C, H // k, W // k), np.float32)
        _lib.check(_lib.lib().cpt_avgpool2d_fwd(f32ptr(x), y.ptr, B, C, H, W, k, stream_ptr()))
        cache.push(x.shape, k)
        return Tensor(y)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor) -> Tensor:
        x_shape, k = cache.pop()
        require_cuda(dy)
        B, C, H, W = x_shape
        dx = DeviceArray.empty(x_shape, np.float32)
        _lib.check(_lib.lib().cpt_avgpool2d_bwd(f32ptr(dy), dx.ptr, B, C, H, W, k, stream_ptr()))
        return Tensor(dx)


def maxpooling2d(x: Tensor, kernel_size: int = 2) -> Tensor:
    return MaxPooling2DFn.forward(PseudoCache(), x, kernel_size)


def avgpooling2d(x: Tensor, kernel_size: int = 2) -> Tensor:
    return AvgPooling2DFn.forward(PseudoCache(), x, kernel_size)
