"""Linear forward/backward.  API of compyute/nn/functional/linear_funcs.py:11-59."""

from __future__ import annotations

from typing import Optional

import numpy as np

from ... import _lib
from ...backend import get_compute_mode
from ...tensors import DeviceArray, ShapeError, Tensor, f32ptr, require_cuda, stream_ptr, workspace
from .activation_funcs import PlainMask
from .functions import Function, FunctionCache, PseudoCache, get_caching_enabled

__all__ = ["linear", "LinearFn"]


class LinearFn(Function):
    """y = x @ w.T (+ b) over the last dim; leading dims are flattened into the GEMM's row dimension, which is
    what the reference's batched matmul + ``.sum(leading)`` computes (linear_funcs.py:15-35)."""

    @staticmethod
    def relu_fusable(x: Tensor, w: Tensor) -> bool:
        """The GEMM epilogue can apply a following ReLU: bf16 mode, 2-D input, Out % 32 == 0 (mask words)."""
        return get_compute_mode() == _lib.MODE_BF16 and x.ndim == 2 and w.shape[0] % 32 == 0

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, w: Tensor, b: Optional[Tensor], relu_cache=None, emit_lp: bool = False) -> Tensor:
        """``relu_cache`` (extension): return ``relu(x @ w.T + b)`` from the GEMM epilogue (``relu_fusable`` must hold); the
        mask is pushed on ``relu_cache`` as ReLUFn.forward would have, ``emit_lp`` also writes the bf16 rows a following Linear
        layer consumes."""
        require_cuda(x, w, b)
        out_f, in_f = w.shape
        if x.shape[-1] != in_f:
            raise ShapeError(f"Expected last input dim {in_f}, got {x.shape[-1]}.")
        n = x.size // in_f
        mode = get_compute_mode()
        L = _lib.lib()
        y = DeviceArray.empty((*x.shape[:-1], out_f), np.float32)
        st = stream_ptr()
        x_bf = w_bf = None
        if mode == _lib.MODE_BF16:
            # operands staged once as bf16 and kept in the cache: x_bf16 is reused by wgrad, w_bf16 by dgrad
            shadow = getattr(x.data, "cl", None) if x.ndim == 2 else None
            w_bf = DeviceArray.empty((L.cpt_cast_bf16_bytes(out_f, in_f),), np.uint8)
            if shadow is not None and shadow[0] == mode:
                x_bf = shadow[1]  # written by the producing ReLU pass
            else:
                x_bf = DeviceArray.empty((L.cpt_cast_bf16_bytes(n, in_f),), np.uint8)
                _lib.check(L.cpt_cast_bf16(f32ptr(x), x_bf.ptr, n, in_f, st))
            _lib.check(L.cpt_cast_bf16(f32ptr(w), w_bf.ptr, out_f, in_f, st))
            if relu_cache is not None:
                want_mask = get_caching_enabled() and not isinstance(relu_cache, PseudoCache)
                mask = DeviceArray.empty(((n * out_f + 31) // 32 * 4,), np.uint8) if want_mask else None
                lp = None
                if emit_lp and out_f % 8 == 0:
                    lp = DeviceArray.empty((L.cpt_cast_bf16_bytes(n, out_f),), np.uint8)
                    y.cl = (mode, lp, None)
                _lib.check(L.cpt_linear_relu_fwd_bf16(x_bf.ptr, w_bf.ptr, f32ptr(b), y.ptr, lp.ptr if lp is not None else None,
                                                      mask.ptr if mask is not None else None, n, in_f, out_f, st))
                relu_cache.push(PlainMask(mask) if mask is not None else None)
            else:
                _lib.check(L.cpt_linear_fwd_bf16(x_bf.ptr, w_bf.ptr, f32ptr(b), y.ptr, n, in_f, out_f, st))
        else:
            ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_FPROP, n, in_f, out_f, mode))
            _lib.check(L.cpt_linear_fwd(f32ptr(x), f32ptr(w), f32ptr(b), y.ptr, n, in_f, out_f, mode, ws, wsb, st))
        cache.push(x, w, b is not None, mode, x_bf, w_bf)
        return Tensor(y)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, dw_out: Optional[DeviceArray] = None, db_out: Optional[DeviceArray] = None,
                 input_relu_mask=None, emit_lp: bool = False) -> tuple[Tensor, Tensor, Optional[Tensor]]:
        """``input_relu_mask`` (extension): the cache entry of the ReLU that produced this layer's input; the returned dx is then
        already that ReLU's dx (``dx * mask``) — from the dgrad epilogue when the mask is a ``PlainMask`` in bf16 mode and
        In % 32 == 0, else by a separate ReLU backward pass; ``emit_lp`` also writes its bf16 rows."""
        x, w, has_bias, mode, x_bf, w_bf = cache.pop()
        require_cuda(dy)
        out_f, in_f = w.shape
        n = x.size // in_f
        L = _lib.lib()
        st = stream_ptr()
        dx = DeviceArray.empty(x.shape, np.float32)
        dw = dw_out.reshape(w.shape) if dw_out is not None else DeviceArray.empty(w.shape, np.float32)
        db = None
        if has_bias:
            db = db_out.reshape((out_f,)) if db_out is not None else DeviceArray.empty((out_f,), np.float32)
        if mode == _lib.MODE_BF16:
            shadow = getattr(dy.data, "cl", None) if dy.ndim == 2 else None
            if shadow is not None and shadow[0] == mode:
                dy_bf = shadow[1]  # written by the ReLU backward pass that produced dy
            else:
                dy_bf = DeviceArray.empty((L.cpt_cast_bf16_bytes(n, out_f),), np.uint8)
                _lib.check(L.cpt_cast_bf16(f32ptr(dy), dy_bf.ptr, n, out_f, st))
            if isinstance(input_relu_mask, PlainMask) and in_f % 32 == 0 and x.ndim == 2:
                lp = None
                if emit_lp and in_f % 8 == 0:
                    lp = DeviceArray.empty((L.cpt_cast_bf16_bytes(n, in_f),), np.uint8)
                    dx.cl = (mode, lp, None)
                _lib.check(L.cpt_linear_dgrad_relu_bf16(dy_bf.ptr, w_bf.ptr, input_relu_mask.bits.ptr, dx.ptr,
                                                        lp.ptr if lp is not None else None, n, in_f, out_f, st))
                input_relu_mask = None  # applied
            else:
                _lib.check(L.cpt_linear_dgrad_bf16(dy_bf.ptr, w_bf.ptr, dx.ptr, n, in_f, out_f, st))
            ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_WGRAD, n, in_f, out_f, mode))
            _lib.check(L.cpt_linear_wgrad_bf16(x_bf.ptr, dy_bf.ptr, dw.ptr, n, in_f, out_f, ws, wsb, st))
            if db is not None:
                _lib.check(L.cpt_channel_sum(f32ptr(dy), db.ptr, n, out_f, 1, ws, wsb, st))
        else:
            ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_DGRAD, n, in_f, out_f, mode))
            _lib.check(L.cpt_linear_dgrad(f32ptr(dy), f32ptr(w), dx.ptr, n, in_f, out_f, mode, ws, wsb, st))
            ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_WGRAD, n, in_f, out_f, mode))
            _lib.check(L.cpt_linear_wgrad(f32ptr(x), f32ptr(dy), dw.ptr, db.ptr if db is not None else None, n, in_f, out_f, mode,
                                          ws, wsb, st))
        dxt = Tensor(dx)
        if input_relu_mask is not None:  # not fused above: the ReLU's own backward pass
            from .activation_funcs import ReLUFn
            tmp = FunctionCache()
            tmp.push(input_relu_mask)
            dxt = ReLUFn.backward(tmp, dxt, emit_lp)
        return dxt, Tensor(dw), (Tensor(db) if db is not None else None)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return LinearFn.forward(PseudoCache(), x, w, b)
