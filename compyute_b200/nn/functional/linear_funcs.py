"""Linear forward/backward.  API of compyute/nn/functional/linear_funcs.py:11-59."""

from __future__ import annotations

from typing import Optional

import numpy as np

from ... import _lib
from ...backend import get_compute_mode
from ...tensors import DeviceArray, ShapeError, Tensor, f32ptr, require_cuda, stream_ptr, workspace
from .functions import Function, FunctionCache, PseudoCache

__all__ = ["linear", "LinearFn"]


class LinearFn(Function):
    """y = x @ w.T (+ b) over the last dim; leading dims are flattened into the GEMM's row dimension, which is
    what the reference's batched matmul + ``.sum(leading)`` computes (linear_funcs.py:15-35)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
        require_cuda(x, w, b)
        out_f, in_f = w.shape
        if x.shape[-1] != in_f:
            raise ShapeError(f"Expected last input dim {in_f}, got {x.shape[-1]}.")
        n = x.size // in_f
        mode = get_compute_mode()
        L = _lib.lib()
        y = DeviceArray.empty((*x.shape[:-1], out_f), np.float32)
        ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_FPROP, n, in_f, out_f, mode))
        _lib.check(L.cpt_linear_fwd(f32ptr(x), f32ptr(w), f32ptr(b), y.ptr, n, in_f, out_f, mode, ws, wsb, stream_ptr()))
        cache.push(x, w, b is not None, mode)
        return Tensor(y)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, dw_out: Optional[DeviceArray] = None,
                 db_out: Optional[DeviceArray] = None) -> tuple[Tensor, Tensor, Optional[Tensor]]:
        x, w, has_bias, mode = cache.pop()
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
        ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_DGRAD, n, in_f, out_f, mode))
        _lib.check(L.cpt_linear_dgrad(f32ptr(dy), f32ptr(w), dx.ptr, n, in_f, out_f, mode, ws, wsb, st))
        ws, wsb = workspace(L.cpt_linear_workspace_size(_lib.OP_WGRAD, n, in_f, out_f, mode))
        _lib.check(L.cpt_linear_wgrad(f32ptr(x), f32ptr(dy), dw.ptr, db.ptr if db is not None else None, n, in_f, out_f, mode,
                                      ws, wsb, st))
        return Tensor(dx), Tensor(dw), (Tensor(db) if db is not None else None)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return LinearFn.forward(PseudoCache(), x, w, b)
