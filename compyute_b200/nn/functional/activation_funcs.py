"""ReLU.  API of compyute/nn/functional/activation_funcs.py:22-51."""

from __future__ import annotations

import numpy as np

from ... import _lib
from ...backend import get_compute_mode
from ...tensors import DeviceArray, Tensor, f32ptr, require_cuda, stream_ptr
from .functions import Function, FunctionCache, PseudoCache, get_caching_enabled

__all__ = ["relu", "ReLUFn", "FUSED_INTO_PRODUCER", "PlainMask"]

# cache marker: this ReLU was evaluated inside its producer (Sequential peephole BatchNorm -> ReLU, see
# normalization_funcs.BatchNormReLU2DFn); the producer's backward already applies dy * (y > 0)
FUSED_INTO_PRODUCER = "fused-into-producer"


class PlainMask:
    """ReLU mask in plain bit order (element e = bit e % 32 of word e / 32), as written by the fused Linear + ReLU epilogue
    (``cpt_linear_relu_fwd_bf16``); ``ReLUFn.backward`` then calls ``cpt_relu_bwd_plain``."""

    __slots__ = ("bits",)

    def __init__(self, bits: DeviceArray) -> None:
        self.bits = bits


def _lp_buffer(a: DeviceArray):
    """bf16 copy of a 2-D (N, C) array for the neighbouring Linear layer, attached as ``a.cl``; only in bf16 mode and when
    C is a multiple of 8 (then the bf16 row pitch of cpt_cast_bf16 equals the row length)."""
    if get_compute_mode() != _lib.MODE_BF16 or a.ndim != 2 or a.shape[1] % 8 != 0:
        return None
    lp = DeviceArray.empty((_lib.lib().cpt_cast_bf16_bytes(a.shape[0], a.shape[1]),), np.uint8)
    a.cl = (_lib.MODE_BF16, lp, None)
    return lp


class ReLUFn(Function):
    """y = max(x, 0); caches the mask ``y > 0`` (:26-34) — bit-packed here (1 bit/element instead of 1 byte)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, emit_lp: bool = False) -> Tensor:
        """``emit_lp`` (extension): also write y as bf16 rows — the pre-cast operand of a Linear layer that consumes it."""
        require_cuda(x)
        n = x.size
        y = DeviceArray.empty(x.shape, np.float32)
        want_mask = get_caching_enabled() and not isinstance(cache, PseudoCache)
        mask = DeviceArray.empty(((n + 31) // 32 * 4,), np.uint8) if want_mask else None
        lp = _lp_buffer(y) if emit_lp else None
        _lib.check(_lib.lib().cpt_relu_fwd_lp(f32ptr(x), y.ptr, mask.ptr if mask is not None else None,
                                              lp.ptr if lp is not None else None, n, stream_ptr()))
        cache.push(mask)
        return Tensor(y)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, emit_lp: bool = False) -> Tensor:
        (mask,) = cache.pop()
        if mask is FUSED_INTO_PRODUCER:
            return dy
        require_cuda(dy)
        dx = DeviceArray.empty(dy.shape, np.float32)
        lp = _lp_buffer(dx) if emit_lp else None
        if isinstance(mask, PlainMask):
            _lib.check(_lib.lib().cpt_relu_bwd_plain(f32ptr(dy), mask.bits.ptr, dx.ptr, lp.ptr if lp is not None else None, dy.size,
                                                     stream_ptr()))
            return Tensor(dx)
        _lib.check(_lib.lib().cpt_relu_bwd_lp(f32ptr(dy), mask.ptr, dx.ptr, lp.ptr if lp is not None else None, dy.size, stream_ptr()))
        return Tensor(dx)


def relu(x: Tensor) -> Tensor:
    return ReLUFn.forward(PseudoCache(), x)
