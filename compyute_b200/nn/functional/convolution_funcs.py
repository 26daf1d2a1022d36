"""Conv2D forward/backward on the B200 kernels.  API of compyute/nn/functional/convolution_funcs.py:218-291.

``Conv2DFn.forward(cache, x, f, b, padding, stride, dilation)`` / ``.backward(cache, dy) -> (dx, df, db|None)``.
The reference evaluates dilate → pad → window view → einsum (and flips/pads again in backward); here each pass
is one implicit-GEMM launch (``cpt_conv2d_fprop/_dgrad/_wgrad``), exact FFMA in "fp32" mode, tcgen05 in
"tf32"/"bf16" mode.  One cache tuple is pushed instead of the reference's four.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from ... import _lib
from ...backend import get_compute_mode
from ...tensors import DeviceArray, ShapeError, Tensor, f32ptr, require_cuda, stream_ptr, workspace
from .functions import Function, FunctionCache, PseudoCache

__all__ = ["conv2d", "Conv2DFn"]

_MODE_PACKED = 100  # cache tag: bf16 layer that ran on the packed-K path (x_cl holds the patch matrix)
_MODE_STRIP = 101   # cache tag: bf16 layer that ran on the strip path (x_cl holds the zero-padded channels-last copy)


def _desc(x_shape, f_shape, padding: int, stride: int, dilation: int) -> _lib.ConvDesc:
    B, Ci, H, W = x_shape
    Co, Ci_f, K, K2 = f_shape
    if Ci_f != Ci:
        raise ShapeError(f"Filter expects {Ci_f} input channels, input has {Ci}.")
    if K != K2:
        raise ShapeError("Only square kernels are supported (like the reference).")
    return _lib.ConvDesc(B, Ci, H, W, Co, K, int(padding), int(stride), int(dilation))


def _out_hw(d: _lib.ConvDesc) -> tuple[int, int]:
    ho, wo = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().cpt_conv2d_out_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)))
    return ho.value, wo.value


class Conv2DFn(Function):
    """2-D cross-correlation, NCHW x OIHW (convolution_funcs.py:218-254)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, f: Tensor, b: Optional[Tensor], padding: int, stride: int,
                dilation: int, emit_stats: bool = False, relu: bool = False) -> Tensor:
        """``emit_stats`` (extension, tensor-core modes): the epilogue also leaves per-channel Σy / Σy² partials on the
        result (``y.data.stats``) for a BatchNorm that consumes it, which then skips its statistics pass.
        ``relu`` (extension; the caller checked ``relu_fusable``): returns ``relu(conv(x))`` from the GEMM epilogue, and
        ``backward`` expects the gradient w.r.t. that — ReLUFn.backward (activation_funcs.py:32-34) is applied while dy is staged,
        with the mask taken from the returned tensor itself, which therefore must not be modified in place."""
        if x.ndim != 4:
            raise ShapeError(f"Expected input to be 4D, got {x.ndim}D.")
        require_cuda(x, f, b)
        L = _lib.lib()
        d = _desc(x.shape, f.shape, padding, stride, dilation)
        ho, wo = _out_hw(d)
        mode = _layer_mode(d)
        if mode == _lib.MODE_FP32X3:
            emit_stats = False  # the exact mode keeps BatchNorm's own two-pass statistics (the epilogue sums are E[a^2] - E[a]^2)
        if relu and not _relu_path_ok(L, d, mode):
            raise NotImplementedError("Conv2DFn.forward(relu=True): layer not covered, check Conv2DFn.relu_fusable first")
        y = DeviceArray.empty((d.B, d.Co, ho, wo), np.float32)
        st = stream_ptr()
        x_cl = None
        if mode == _lib.MODE_FP32:
            ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_FPROP, ctypes.byref(d), mode))
            _lib.check(L.cpt_conv2d_fprop(ctypes.byref(d), f32ptr(x), f32ptr(f), f32ptr(b), y.ptr, mode, ws, wsb, st))
        elif L.cpt_conv2d_packed_bytes(ctypes.byref(d), mode):
            # first layers (tiny Ci): explicit bf16 patch matrix with K = Ci*k*k packed densely, kept for wgrad
            x_cl = DeviceArray.empty((L.cpt_conv2d_packed_bytes(ctypes.byref(d), mode),), np.uint8)
            _lib.check(L.cpt_conv2d_im2col_pack(ctypes.byref(d), f32ptr(x), x_cl.ptr, st))
            ws, wsb = workspace(L.cpt_conv2d_packed_workspace_size(_lib.OP_FPROP, ctypes.byref(d)))
            if emit_stats:
                stats = DeviceArray.empty((L.cpt_conv2d_stats_bytes(ctypes.byref(d)),), np.uint8)
                _lib.check(L.cpt_conv2d_fprop_packed_stats(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr, stats.ptr, ws, wsb, st))
                y.stats = (stats, L.cpt_conv2d_stats_slots(), b)
            elif relu:
                _lib.check(L.cpt_conv2d_fprop_packed_relu(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr, ws, wsb, st))
            else:
                _lib.check(L.cpt_conv2d_fprop_packed(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr, ws, wsb, st))
            mode = _MODE_PACKED
        elif not relu and L.cpt_conv2d_strip_supported(ctypes.byref(d), mode) and not _has_dense_shadow(x, mode):
            # stride-1 same-padded small-channel layer: zero-padded channels-last operand, one strip per 128 outputs
            shadow = getattr(x.data, "cl", None)
            if shadow is not None and shadow[0] == (_MODE_STRIP, d.pad):
                x_cl = shadow[1]  # the producer already wrote the padded operand
            else:
                x_cl = DeviceArray.empty((L.cpt_channels_last_padded_bytes(d.B, d.Ci, d.H, d.W, d.pad),), np.uint8)
                _lib.check(L.cpt_to_channels_last_padded(f32ptr(x), x_cl.ptr, d.B, d.Ci, d.H, d.W, d.pad, None, None, 0, st))
            ws, wsb = workspace(L.cpt_conv2d_strip_workspace_size(_lib.OP_FPROP, ctypes.byref(d)))
            stats = None
            if emit_stats:
                stats = DeviceArray.empty((L.cpt_conv2d_stats_bytes(ctypes.byref(d)),), np.uint8)
                y.stats = (stats, L.cpt_conv2d_stats_slots(), b)
            _lib.check(L.cpt_conv2d_fprop_strip(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr,
                                                stats.ptr if stats is not None else None, ws, wsb, st))
            mode = _MODE_STRIP
        else:
            shadow = getattr(x.data, "cl", None)
            if shadow is not None and shadow[0] == mode:
                x_cl = shadow[1]  # the producer (BatchNorm/ReLU pass) already wrote the channels-last operand
            else:
                # stage x once as channels-last; kept in the cache so wgrad does not convert it again
                x_cl = DeviceArray.empty((L.cpt_channels_last_bytes(d.B, d.Ci, d.H, d.W, mode),), np.uint8)
                _lib.check(L.cpt_to_channels_last(f32ptr(x), x_cl.ptr, d.B, d.Ci, d.H, d.W, mode, None, None, 0, st))
            ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_FPROP, ctypes.byref(d), mode))
            if emit_stats:
                stats = DeviceArray.empty((L.cpt_conv2d_stats_bytes(ctypes.byref(d)),), np.uint8)
                _lib.check(L.cpt_conv2d_fprop_cl_stats(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr, stats.ptr, mode, ws, wsb, st))
                y.stats = (stats, L.cpt_conv2d_stats_slots(), b)
            elif relu:
                _lib.check(L.cpt_conv2d_fprop_cl_relu(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr, mode, ws, wsb, st))
            else:
                _lib.check(L.cpt_conv2d_fprop_cl(ctypes.byref(d), x_cl.ptr, f32ptr(f), f32ptr(b), y.ptr, mode, ws, wsb, st))
        cache.push(x, f, b is not None, d, mode, x_cl, y if relu else None)
        return Tensor(y)

    @staticmethod
    def relu_fusable(x: Tensor, f: Tensor, padding: int, stride: int, dilation: int) -> bool:
        """Can ``forward(..., relu=True)`` serve this layer?  Tensor-core modes on the dense / packed paths, with the input
        gradient on the tensor-core path as well (the exact fallbacks read dy unstaged)."""
        if x.ndim != 4 or not isinstance(x.data, DeviceArray) or f.ndim != 4 or f.shape[1] != x.shape[1] or f.shape[2] != f.shape[3]:
            return False
        d = _desc(x.shape, f.shape, padding, stride, dilation)
        return _relu_path_ok(_lib.lib(), d, _layer_mode(d))

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, df_out: Optional[DeviceArray] = None,
                 db_out: Optional[DeviceArray] = None) -> tuple[Tensor, Tensor, Optional[Tensor]]:
        """``df_out``/``db_out``: optional preallocated gradient slots (flat DP arena); extension of the reference API."""
        x, f, has_bias, d, mode, x_cl, gate = cache.pop()
        require_cuda(dy)
        L = _lib.lib()
        st = stream_ptr()
        dref = ctypes.byref(d)
        dx = DeviceArray.empty(x.shape, np.float32)
        df = df_out.reshape(f.shape) if df_out is not None else DeviceArray.empty(f.shape, np.float32)
        db = None
        if has_bias:
            db = db_out.reshape((d.Co,)) if db_out is not None else DeviceArray.empty((d.Co,), np.float32)
        dbp = db.ptr if db is not None else None
        ho, wo = dy.shape[2], dy.shape[3]
        if mode == _MODE_PACKED:
            dy_cl = _staged_dy(L, dy, d, ho, wo, _lib.MODE_BF16, db, st, gate)
            ws, wsb = workspace(L.cpt_conv2d_packed_workspace_size(_lib.OP_DGRAD, dref))
            _lib.check(L.cpt_conv2d_dgrad_packed(dref, dy_cl.ptr, f32ptr(f), dx.ptr, ws, wsb, st))
            ws, wsb = workspace(L.cpt_conv2d_packed_workspace_size(_lib.OP_WGRAD, dref))
            _lib.check(L.cpt_conv2d_wgrad_packed(dref, x_cl.ptr, dy_cl.ptr, df.ptr, ws, wsb, st))
            return Tensor(dx), Tensor(df), (Tensor(db) if db is not None else None)
        if mode == _MODE_STRIP:
            dy_pad = _staged_dy_padded(L, dy, d, db, st)
            ws, wsb = workspace(L.cpt_conv2d_strip_workspace_size(_lib.OP_DGRAD, dref))
            _lib.check(L.cpt_conv2d_dgrad_strip(dref, dy_pad.ptr, f32ptr(f), dx.ptr, ws, wsb, st))
            ws, wsb = workspace(L.cpt_conv2d_strip_workspace_size(_lib.OP_WGRAD, dref))
            _lib.check(L.cpt_conv2d_wgrad_padded(dref, x_cl.ptr, dy_pad.ptr, df.ptr, ws, wsb, st))
            return Tensor(dx), Tensor(df), (Tensor(db) if db is not None else None)
        tc_dgrad = mode != _lib.MODE_FP32 and bool(L.cpt_conv2d_dgrad_cl_supported(dref, mode))
        if mode == _lib.MODE_FP32:
            ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_DGRAD, dref, mode))
            _lib.check(L.cpt_conv2d_dgrad(dref, f32ptr(dy), f32ptr(f), dx.ptr, mode, ws, wsb, st))
            ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_WGRAD, dref, mode))
            _lib.check(L.cpt_conv2d_wgrad(dref, f32ptr(x), f32ptr(dy), df.ptr, dbp, mode, ws, wsb, st))
        else:
            # dy staged once (db fused into the staging pass), shared by dgrad and wgrad
            dy_cl = _staged_dy(L, dy, d, ho, wo, mode, db, st, gate)
            if tc_dgrad:
                ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_DGRAD, dref, mode))
                _lib.check(L.cpt_conv2d_dgrad_cl(dref, dy_cl.ptr, f32ptr(f), dx.ptr, mode, ws, wsb, st))
            else:  # geometry outside the TMA im2col limits: exact path
                ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_DGRAD, dref, _lib.MODE_FP32))
                _lib.check(L.cpt_conv2d_dgrad(dref, f32ptr(dy), f32ptr(f), dx.ptr, _lib.MODE_FP32, ws, wsb, st))
            ws, wsb = workspace(L.cpt_conv2d_workspace_size(_lib.OP_WGRAD, dref, mode))
            _lib.check(L.cpt_conv2d_wgrad_cl(dref, x_cl.ptr, dy_cl.ptr, df.ptr, mode, ws, wsb, st))
        return Tensor(dx), Tensor(df), (Tensor(db) if db is not None else None)


def _layer_mode(d) -> int:
    """Compute mode of one layer: the global mode, or the exact FFMA kernels where the tensor-core path does not apply."""
    mode = get_compute_mode()
    if mode != _lib.MODE_FP32 and (d.K * d.K > 64 or (d.K - 1) * d.dil > 255 or d.pad > 127 or d.stride > 8):
        return _lib.MODE_FP32  # outside the TMA im2col limits (8-bit corners, 64 taps): exact path for this layer
    if mode == _lib.MODE_FP32X3 and d.Ci < 8:
        return _lib.MODE_FP32  # a tap's k-slice is 32 channels wide: with < 8 real channels the FFMA kernel is the faster exact path
    return mode


def _relu_path_ok(L, d, mode: int) -> bool:
    if mode == _lib.MODE_FP32:
        return False
    dref = ctypes.byref(d)
    return bool(L.cpt_conv2d_packed_bytes(dref, mode)) or bool(L.cpt_conv2d_dgrad_cl_supported(dref, mode))


def _staged_dy(L, dy: Tensor, d, ho: int, wo: int, mode: int, db: Optional[DeviceArray], st,
               gate: Optional[DeviceArray] = None) -> DeviceArray:
    """Channels-last copy of dy (+ db = dy.sum((0, 2, 3)), convolution_funcs.py:252).  When dy was produced by a
    BatchNorm backward pass that already emitted it (``dy.data.cl``), nothing is staged.  ``gate``: the layer ran with its ReLU
    in the epilogue — dy is the gradient behind that ReLU, and ``dy * (gate > 0)`` is what gets staged and summed."""
    if gate is not None:
        dy_cl = DeviceArray.empty((L.cpt_channels_last_bytes(d.B, d.Co, ho, wo, mode),), np.uint8)
        ws, wsb = workspace(L.cpt_to_channels_last_workspace_size(d.B, d.Co, ho, wo))
        _lib.check(L.cpt_to_channels_last_gated(f32ptr(dy), gate.ptr, dy_cl.ptr, d.B, d.Co, ho, wo, mode,
                                                db.ptr if db is not None else None, ws, wsb, st))
        return dy_cl
    shadow = getattr(dy.data, "cl", None)
    if shadow is not None and shadow[0] == mode:
        if db is not None:
            if shadow[2] is not None:
                db.copy_from(shadow[2])
            else:
                ws, wsb = workspace(d.Co * 64 * 4)
                _lib.check(L.cpt_channel_sum(f32ptr(dy), db.ptr, d.B, d.Co, ho * wo, ws, wsb, st))
        return shadow[1]
    dy_cl = DeviceArray.empty((L.cpt_channels_last_bytes(d.B, d.Co, ho, wo, mode),), np.uint8)
    ws, wsb = workspace(L.cpt_to_channels_last_workspace_size(d.B, d.Co, ho, wo))
    _lib.check(L.cpt_to_channels_last(f32ptr(dy), dy_cl.ptr, d.B, d.Co, ho, wo, mode, db.ptr if db is not None else None, ws, wsb, st))
    return dy_cl


def _has_dense_shadow(x: Tensor, mode: int) -> bool:
    """The producer already wrote the dense channels-last operand of the im2col path: using it beats re-staging x padded."""
    shadow = getattr(x.data, "cl", None)
    return shadow is not None and shadow[0] == mode


def _staged_dy_padded(L, dy: Tensor, d, db: Optional[DeviceArray], st) -> DeviceArray:
    """Zero-padded channels-last copy of dy for the strip path (+ db fused into the staging pass)."""
    shadow = getattr(dy.data, "cl", None)
    if shadow is not None and shadow[0] == (_MODE_STRIP, d.pad):
        if db is not None:
            if shadow[2] is not None:
                db.copy_from(shadow[2])
            else:
                ws, wsb = workspace(d.Co * 64 * 4)
                _lib.check(L.cpt_channel_sum(f32ptr(dy), db.ptr, d.B, d.Co, d.H * d.W, ws, wsb, st))
        return shadow[1]
    dy_pad = DeviceArray.empty((L.cpt_channels_last_padded_bytes(d.B, d.Co, d.H, d.W, d.pad),), np.uint8)
    ws, wsb = workspace(L.cpt_to_channels_last_workspace_size(d.B, d.Co, d.H, d.W))
    _lib.check(L.cpt_to_channels_last_padded(f32ptr(dy), dy_pad.ptr, d.B, d.Co, d.H, d.W, d.pad, db.ptr if db is not None else None,
                                             ws, wsb, st))
    return dy_pad


def conv2d(x: Tensor, f: Tensor, b: Optional[Tensor] = None, padding: int = 0, stride: int = 1,
           dilation: int = 1) -> Tensor:
    """User-level wrapper (convolution_funcs.py:257-291)."""
    return Conv2DFn.forward(PseudoCache(), x, f, b, padding, stride, dilation)
