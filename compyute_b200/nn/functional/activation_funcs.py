"""ReLU.  API of compyute/nn/functional/activation_funcs.py:22-51."""

from __future__ import annotations

import numpy as np

from ... import _lib
from ...tensors import DeviceArray, Tensor, f32ptr, require_cuda, stream_ptr
from .functions import Function, FunctionCache, PseudoCache, get_caching_enabled

__all__ = ["relu", "ReLUFn", "FUSED_INTO_PRODUCER"]

# cache marker: this ReLU was evaluated inside its producer (Sequential peephole BatchNorm -> ReLU, see
# normalization_funcs.BatchNormReLU2DFn); the producer's backward already applies dy * (y > 0)
FUSED_INTO_PRODUCER = "fused-into-producer"


class ReLUFn(Function):
    """y = max(x, 0); caches the mask ``y > 0`` (:26-34) — bit-packed here (1 bit/element instead of 1 byte)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor) -> Tensor:
        require_cuda(x)
        n = x.size
        y = DeviceArray.empty(x.shape, np.float32)
        want_mask = get_caching_enabled() and not isinstance(cache, PseudoCache)
        mask = DeviceArray.empty(((n + 31) // 32 * 4,), np.uint8) if want_mask else None
        _lib.check(_lib.lib().cpt_relu_fwd(f32ptr(x), y.ptr, mask.ptr if mask is not None else None, n, stream_ptr()))
        cache.push(mask)
        return Tensor(y)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor) -> Tensor:
        (mask,) = cache.pop()
        if mask is FUSED_INTO_PRODUCER:
            return dy
        require_cuda(dy)
        dx = DeviceArray.empty(dy.shape, np.float32)
        _lib.check(_lib.lib().cpt_relu_bwd(f32ptr(dy), mask.ptr, dx.ptr, dy.size, stream_ptr()))
        return Tensor(dx)


def relu(x: Tensor) -> Tensor:
    return ReLUFn.forward(PseudoCache(), x)
