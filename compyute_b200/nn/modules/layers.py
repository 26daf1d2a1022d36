"""Layers of the CNN hot path with the reference's constructor arguments, attribute names (``w``, ``b``,
``rmean``, ``rvar``) and state-dict keys:

  Conv2D        compyute/nn/modules/convolutions.py:118-204
  Linear        compyute/nn/modules/linear.py:15-72
  MaxPooling2D  compyute/nn/modules/poolings.py:43-63,  AvgPooling2D  poolings.py:15-40
  BatchNorm1D   compyute/nn/modules/normalizations.py:20-93,  BatchNorm2D  normalizations.py:96-171
  ReLU          compyute/nn/modules/activations.py:101-120
  Flatten       compyute/nn/modules/shapes.py:12-27
  Dropout       compyute/nn/modules/regularizations.py:12-34
"""

from __future__ import annotations

import math
from typing import Literal, Optional

import numpy as np

from ...backend import select_device
from ...tensors import Tensor, tensor
from ..functional.activation_funcs import ReLUFn
from ..functional.convolution_funcs import Conv2DFn
from ..functional.linear_funcs import LinearFn
from ..functional.normalization_funcs import BatchNorm1DFn, BatchNorm2DFn, BatchNormReLU1DFn, BatchNormReLU2DFn
from ..functional.pooling_funcs import AvgPooling2DFn, MaxPooling2DFn
from ..functional.regularization_funcs import DropoutFn
from ..functional.shape_funcs import FlattenFn
from ..parameter import Buffer, Parameter
from .module import Module

__all__ = ["Conv2D", "Linear", "MaxPooling2D", "AvgPooling2D", "BatchNorm1D", "BatchNorm2D", "ReLU", "Flatten", "Dropout"]

PaddingLike = int | Literal["valid", "same"]


def _uniform(shape, k: float) -> Tensor:
    """U(-k, k) from NumPy's legacy global stream, like compyute/random/random.py:149 — same seed, same weights."""
    return tensor(np.random.uniform(-k, k, shape).astype(np.float32), device=select_device(None))


def _const(shape, value: float) -> Tensor:
    return tensor(np.full(shape, value, dtype=np.float32), device=select_device(None))


class Conv2D(Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, padding: PaddingLike = "valid", stride: int = 1,
                 dilation: int = 1, bias: bool = True, label: Optional[str] = None) -> None:
        super().__init__(label)
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        if isinstance(padding, int):
            self.padding = padding
        else:  # convolutions.py:23-28: "same" is the symmetric (k*d - 1) // 2
            self.padding = 0 if padding == "valid" else (kernel_size * dilation - 1) // 2
        self.stride, self.dilation, self.bias = stride, dilation, bias
        k = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
        self.w = Parameter(_uniform((out_channels, in_channels, kernel_size, kernel_size), k))
        self.b = Parameter(_uniform((out_channels,), k)) if bias else None

    _emit_stats = False  # hint set by the enclosing Sequential: a BatchNorm2D consumes the output

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return Conv2DFn.forward(self.fcache, x, self.w, self.b, self.padding, self.stride, self.dilation,
                                self._emit_stats and self._is_training)

    def forward_relu(self, x: Tensor) -> Tensor:
        """``relu(self(x))`` from the convolution's epilogue (called by ``Sequential`` for Conv2D -> ReLU, which also marks the
        ReLU's cache so that its backward passes dy through): the ReLU's backward is folded into this layer's staging of dy."""
        return Conv2DFn.forward(self.fcache, x, self.w, self.b, self.padding, self.stride, self.dilation, False, relu=True)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        dx, dw, db = Conv2DFn.backward(self.fcache, dy, self.grad_slot(self.w), self.grad_slot(self.b))
        self.update_parameter_grad(self.w, dw)
        self.update_parameter_grad(self.b, db)
        return dx


class Linear(Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, label: Optional[str] = None) -> None:
        super().__init__(label)
        self.in_channels, self.out_channels, self.bias = in_channels, out_channels, bias
        k = 1.0 / math.sqrt(in_channels)
        self.w = Parameter(_uniform((out_channels, in_channels), k))
        self.b = Parameter(_uniform((out_channels,), k)) if bias else None

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return LinearFn.forward(self.fcache, x, self.w, self.b)

    def forward_relu(self, x: Tensor, relu: Module) -> Tensor:
        """``relu(self(x))`` from the GEMM epilogue (bf16 mode; called by ``Sequential`` for Linear -> ReLU): the ReLU mask goes to
        ``relu``'s cache, so both backward methods run unchanged; the bf16 rows for a following Linear are written as well when
        the enclosing Sequential planned that hint for the ReLU."""
        return LinearFn.forward(self.fcache, x, self.w, self.b, relu.fcache, relu._emit_lp_fwd)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        dx, dw, db = LinearFn.backward(self.fcache, dy, self.grad_slot(self.w), self.grad_slot(self.b))
        self.update_parameter_grad(self.w, dw)
        self.update_parameter_grad(self.b, db)
        return dx

    def backward_relu(self, dy: Tensor, relu: Module) -> Tensor:
        """``relu.backward(self.backward(dy))`` with the ReLU's ``dx * mask`` applied in the dgrad epilogue (``relu`` produced
        this layer's input and holds a plain-order mask; called by ``Sequential.backward``)."""
        (mask,) = relu.fcache.pop()
        dx, dw, db = LinearFn.backward(self.fcache, dy, self.grad_slot(self.w), self.grad_slot(self.b), mask, relu._emit_lp_bwd)
        self.update_parameter_grad(self.w, dw)
        self.update_parameter_grad(self.b, db)
        return dx


class MaxPooling2D(Module):
    def __init__(self, kernel_size: int = 2, label: Optional[str] = None) -> None:
        super().__init__(label)
        self.kernel_size = kernel_size

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return MaxPooling2DFn.forward(self.fcache, x, self.kernel_size)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        return MaxPooling2DFn.backward(self.fcache, dy)


class AvgPooling2D(Module):
    def __init__(self, kernel_size: int = 2, label: Optional[str] = None) -> None:
        super().__init__(label)
        self.kernel_size = kernel_size

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return AvgPooling2DFn.forward(self.fcache, x, self.kernel_size)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        return AvgPooling2DFn.backward(self.fcache, dy)


class _BatchNorm(Module):
    _fn = None
    _fn_relu = None  # fused BatchNorm -> ReLU (Sequential peephole)
    # producer-side staging hints set by the enclosing Sequential: the consumer of the output is a tensor-core convolution
    # (forward), the producer of the input is one (backward; with_sum: it has a bias whose gradient is dx's channel sum)
    _emit_cl_fwd = False
    _emit_cl_bwd = False
    _emit_cl_bwd_sum = False

    def __init__(self, channels: int, eps: float = 1e-5, m: float = 0.1, label: Optional[str] = None) -> None:
        super().__init__(label)
        self.channels, self.eps, self.m = channels, eps, m
        self.w = Parameter(_const((channels,), 1.0))
        self.b = Parameter(_const((channels,), 0.0))
        self.rmean = Buffer(_const((channels,), 0.0))
        self.rvar = Buffer(_const((channels,), 1.0))

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return self._forward(self._fn, x)

    def forward_relu(self, x: Tensor) -> Tensor:
        """``relu(self(x))`` in one pass; ``backward`` then expects the gradient w.r.t. the ReLU output (the cache entry
        records the fusion).  Called by ``Sequential`` for a BatchNorm directly followed by a ReLU."""
        return self._forward(self._fn_relu, x)

    def forward_add_relu(self, x: Tensor, skip: Tensor, relu: Module) -> Tensor:
        """``relu(self(x) + skip)`` in one pass (residual tail): this module's cache entry is the plain BatchNorm one, the
        ReLU mask is pushed on ``relu``'s cache, so both backward methods run unchanged."""
        y, rmean, rvar = self._fn.forward(self.fcache, x, self.rmean, self.rvar, self.w, self.b, self.m, self.eps, self._is_training,
                                          False, skip, relu.fcache)
        self.rmean.data = rmean.data
        self.rvar.data = rvar.data
        return y

    def forward_relu_pool2(self, x: Tensor) -> Tensor:
        """``maxpool2(relu(self(x)))`` in one pass; ``backward`` then expects the gradient of the pooled output (the cache entry
        records the fusion).  Called by ``Sequential`` for BatchNorm2D -> ReLU -> MaxPooling2D(2)."""
        y, rmean, rvar = self._fn.forward(self.fcache, x, self.rmean, self.rvar, self.w, self.b, self.m, self.eps, self._is_training,
                                          False, None, None, True)
        self.rmean.data = rmean.data
        self.rvar.data = rvar.data
        return y

    def _forward(self, fn, x: Tensor) -> Tensor:
        extra = (True,) if (self._emit_cl_fwd and x.ndim == 4) else ()
        y, rmean, rvar = fn.forward(self.fcache, x, self.rmean, self.rvar, self.w, self.b, self.m, self.eps, self._is_training,
                                    *extra)
        self.rmean.data = rmean.data  # rebinding, like normalizations.py:163-164
        self.rvar.data = rvar.data
        return y

    def backward(self, dy: Tensor) -> Tensor:  # not wrapped in the reference either (normalizations.py:89, 167)
        extra = (True, self._emit_cl_bwd_sum) if (self._emit_cl_bwd and dy.ndim == 4) else ()
        dx, dw, db = self._fn.backward(self.fcache, dy, self.grad_slot(self.w), self.grad_slot(self.b), *extra)
        self.update_parameter_grad(self.w, dw)
        self.update_parameter_grad(self.b, db)
        return dx


class BatchNorm1D(_BatchNorm):
    _fn = BatchNorm1DFn
    _fn_relu = BatchNormReLU1DFn


class BatchNorm2D(_BatchNorm):
    _fn = BatchNorm2DFn
    _fn_relu = BatchNormReLU2DFn


class ReLU(Module):
    # hints set by the enclosing Sequential: a Linear layer consumes the output (forward) / produced the input (backward),
    # so the pass also writes the bf16 operand that layer would otherwise cast
    _emit_lp_fwd = False
    _emit_lp_bwd = False

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return ReLUFn.forward(self.fcache, x, self._emit_lp_fwd and x.ndim == 2)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        return ReLUFn.backward(self.fcache, dy, self._emit_lp_bwd and dy.ndim == 2)


class Flatten(Module):
    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return FlattenFn.forward(self.fcache, x)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        return FlattenFn.backward(self.fcache, dy)


class Dropout(Module):
    def __init__(self, p: float = 0.5, label: Optional[str] = None) -> None:
        super().__init__(label)
        self.p = p

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return DropoutFn.forward(self.fcache, x, self.p, self._is_training)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        return DropoutFn.backward(self.fcache, dy)
