"""BatchNorm 1-D / 2-D.  API of compyute/nn/functional/normalization_funcs.py:10-224."""

from __future__ import annotations

import numpy as np

from ... import _lib, distributed, graph
from ...backend import get_compute_mode
from ...tensors import DeviceArray, ShapeError, Tensor, f32ptr, require_cuda, stream_ptr, workspace
from .functions import Function, FunctionCache, PseudoCache, get_caching_enabled

__all__ = ["batchnorm1d", "batchnorm2d", "BatchNorm1DFn", "BatchNorm2DFn", "BatchNormReLU1DFn", "BatchNormReLU2DFn",
           "residual_tail_supported", "pool2_fusion_supported"]


ACT_NONE, ACT_RELU = 0, 1  # CPT_ACT_*
ACT_RELU_POOL2 = 2          # cache marker only: BatchNorm2D -> ReLU -> MaxPooling2D(2) evaluated as one pass (cpt_bn_relu_pool2_*)


def _cl_ok(x) -> bool:
    """The producer-side channels-last copy is for the bf16 tensor-core convolutions and 4-D activations only."""
    return x.ndim == 4 and get_compute_mode() == _lib.MODE_BF16


def residual_tail_supported(x) -> bool:
    """cpt_bn_add_relu_apply: 4-D activations with H*W % 4 == 0 and fewer than 2^31 elements."""
    return x.ndim == 4 and (x.shape[2] * x.shape[3]) % 4 == 0 and x.size < (1 << 31)


def pool2_fusion_supported(x) -> bool:
    """cpt_bn_relu_pool2_*: 4-D activations with H even and W % 4 == 0; per-shard statistics only (the synchronised backward
    sums do not know the pooling mask)."""
    return x.ndim == 4 and x.shape[2] % 2 == 0 and x.shape[3] % 4 == 0 and not distributed.sync_batchnorm_active()


def _bn_forward(cache, x, rmean, rvar, w, b, m, eps, training, N, C, HW, act=ACT_NONE, emit_cl=False, skip=None, relu_cache=None,
                pool2=False):
    """``skip`` / ``relu_cache`` (extension): evaluate ``relu(bn(x) + skip)`` — the tail of a residual block — in one pass; the
    ReLU mask goes to ``relu_cache`` exactly as ReLUFn.forward would have pushed it, the BatchNorm cache entry is the plain one
    (its backward receives the ReLU's dx).  The statistics step runs with y == NULL, then cpt_bn_add_relu_apply applies."""
    require_cuda(x, rmean, rvar, w, b)
    L = _lib.lib()
    st = stream_ptr()
    if pool2:  # relu(bn(x)) max-pooled 2x2 in the same pass: only the pooled tensor exists
        act, emit_cl = ACT_NONE, False
        y = DeviceArray.empty((x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2), np.float32)
    else:
        y = DeviceArray.empty(x.shape, np.float32)
    y_stats = y.ptr if (skip is None and not pool2) else None  # NULL: statistics only
    y_cl = sync = None
    if emit_cl and _cl_ok(x):  # the consumer is a tensor-core convolution: write its bf16 NHWC operand in the same pass
        y_cl = DeviceArray.empty((L.cpt_channels_last_bytes(N, C, HW, 1, _lib.MODE_BF16),), np.uint8)
        y.cl = (_lib.MODE_BF16, y_cl, None)
    save_mean = DeviceArray.empty((C,), np.float32)
    save_rstd = DeviceArray.empty((C,), np.float32)
    if training:
        if graph.is_capturing():  # running stats must chain across replays: update the existing buffers in place
            new_rmean, new_rvar = rmean.data, rvar.data
        else:
            new_rmean = DeviceArray.empty((C,), np.float32)
            new_rvar = DeviceArray.empty((C,), np.float32)
        pre = getattr(x.data, "stats", None)
        ws, wsb = workspace(L.cpt_bn_workspace_size(N, C, HW))
        if distributed.sync_batchnorm_active():  # statistics over the global batch: local (mean, M2, n) -> all-gather -> merge
            stats = DeviceArray.empty((3, C), np.float32)
            _lib.check(L.cpt_bn_local_stats(f32ptr(x), stats.ptr, N, C, HW, ws, wsb, st))
            gathered = distributed.all_gather(stats)
            sync = DeviceArray.empty((1,), np.float32)  # global element count, stays on the device for backward
            _lib.check(L.cpt_bn_act_fwd_train_merged(f32ptr(x), f32ptr(w), f32ptr(b), f32ptr(rmean), f32ptr(rvar), y_stats,
                                                     y_cl.ptr if y_cl is not None else None, new_rmean.ptr, new_rvar.ptr,
                                                     save_mean.ptr, save_rstd.ptr, N, C, HW, float(m), float(eps), act,
                                                     gathered.ptr, gathered.shape[0], sync.ptr, st))
        elif pre is not None:  # the producing convolution's epilogue already summed the batch statistics
            _lib.check(L.cpt_bn_act_fwd_train_presum(f32ptr(x), f32ptr(w), f32ptr(b), f32ptr(rmean), f32ptr(rvar), y_stats,
                                                     y_cl.ptr if y_cl is not None else None, new_rmean.ptr, new_rvar.ptr,
                                                     save_mean.ptr, save_rstd.ptr, N, C, HW, float(m), float(eps), act,
                                                     pre[0].ptr, pre[1], f32ptr(pre[2]), st))
        elif y_cl is not None:
            _lib.check(L.cpt_bn_act_fwd_train_cl(f32ptr(x), f32ptr(w), f32ptr(b), f32ptr(rmean), f32ptr(rvar), y.ptr, y_cl.ptr,
                                                 new_rmean.ptr, new_rvar.ptr, save_mean.ptr, save_rstd.ptr, N, C, HW, float(m),
                                                 float(eps), act, ws, wsb, st))
        else:
            _lib.check(L.cpt_bn_act_fwd_train(f32ptr(x), f32ptr(w), f32ptr(b), f32ptr(rmean), f32ptr(rvar), y_stats, new_rmean.ptr,
                                              new_rvar.ptr, save_mean.ptr, save_rstd.ptr, N, C, HW, float(m), float(eps), act, ws,
                                              wsb, st))
        rmean, rvar = Tensor(new_rmean), Tensor(new_rvar)
    elif y_cl is not None:
        _lib.check(L.cpt_bn_act_fwd_eval_cl(f32ptr(x), f32ptr(w), f32ptr(b), f32ptr(rmean), f32ptr(rvar), y.ptr, y_cl.ptr,
                                            save_mean.ptr, save_rstd.ptr, N, C, HW, float(eps), act, st))
    else:
        _lib.check(L.cpt_bn_act_fwd_eval(f32ptr(x), f32ptr(w), f32ptr(b), f32ptr(rmean), f32ptr(rvar), y_stats, save_mean.ptr,
                                         save_rstd.ptr, N, C, HW, float(eps), act, st))
    if pool2:
        _lib.check(L.cpt_bn_relu_pool2_fwd(f32ptr(x), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, y.ptr, x.shape[0], C,
                                           x.shape[2], x.shape[3], st))
        cache.push(x, w, b, save_mean, save_rstd, (N, C, HW), ACT_RELU_POOL2, None)
        return Tensor(y), rmean, rvar
    if skip is not None:
        require_cuda(skip)
        if skip.shape != x.shape:
            raise ShapeError(f"residual shapes {x.shape} and {skip.shape} differ")
        want_mask = relu_cache is not None and get_caching_enabled() and not isinstance(relu_cache, PseudoCache)
        mask = DeviceArray.empty(((x.size + 31) // 32 * 4,), np.uint8) if want_mask else None
        _lib.check(L.cpt_bn_add_relu_apply(f32ptr(x), f32ptr(skip), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, y.ptr,
                                           mask.ptr if mask is not None else None, N, C, HW, st))
        if relu_cache is not None:
            relu_cache.push(mask)
    # the reference caches (w, dims, std, x_norm); x_norm is recomputed from (x, mean, rstd) in backward instead
    # with a fused ReLU the mask is recomputed from (x, mean, rstd, w, b) in backward: nothing extra is cached but b
    cache.push(x, w, b if act else None, save_mean, save_rstd, (N, C, HW), act, sync)
    return Tensor(y), rmean, rvar


def _bn_backward(cache, dy, dw_out=None, db_out=None, emit_cl=False, emit_sum=False):
    """``emit_cl``: also write dx as channels-last bf16 (the dy operand of the producing convolution's backward);
    ``emit_sum``: and its per-channel sums (that convolution's bias gradient)."""
    x, w, b, save_mean, save_rstd, (N, C, HW), act, sync = cache.pop()
    require_cuda(dy)
    L = _lib.lib()
    dx = DeviceArray.empty(x.shape, np.float32)
    dw = dw_out.reshape((C,)) if dw_out is not None else DeviceArray.empty((C,), np.float32)
    db = db_out.reshape((C,)) if db_out is not None else DeviceArray.empty((C,), np.float32)
    if act == ACT_RELU_POOL2:  # dy is the gradient of the POOLED output; tie mask and ReLU mask are recomputed from x
        ws, wsb = workspace(L.cpt_bn_workspace_size(N, C, HW))
        _lib.check(L.cpt_bn_relu_pool2_bwd(f32ptr(x), f32ptr(dy), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, dx.ptr, dw.ptr,
                                           db.ptr, x.shape[0], C, x.shape[2], x.shape[3], ws, wsb, stream_ptr()))
        return Tensor(dx), Tensor(dw), Tensor(db)
    if sync is not None:  # forward took global statistics: the two backward sums are global too
        want_cl = emit_cl and _cl_ok(x)
        dx_cl = DeviceArray.empty((L.cpt_channels_last_bytes(N, C, HW, 1, _lib.MODE_BF16),), np.uint8) if want_cl else None
        csum = DeviceArray.empty((C,), np.float32) if want_cl and emit_sum else None
        ws, wsb = workspace(L.cpt_bn_cl_workspace_size(N, C, HW) if want_cl else L.cpt_bn_workspace_size(N, C, HW))
        sums = DeviceArray.empty((2, C), np.float32)
        _lib.check(L.cpt_bn_act_bwd_local_sums(f32ptr(x), f32ptr(dy), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, sums.ptr,
                                               dw.ptr, db.ptr, N, C, HW, act, ws, wsb, stream_ptr()))
        distributed.all_reduce_sum(sums)
        _lib.check(L.cpt_bn_act_bwd_apply_global(f32ptr(x), f32ptr(dy), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, sums.ptr,
                                                 sync.ptr, dx.ptr, dx_cl.ptr if want_cl else None,
                                                 csum.ptr if csum is not None else None, N, C, HW, act, ws, wsb, stream_ptr()))
        if want_cl:
            dx.cl = (_lib.MODE_BF16, dx_cl, csum)
        return Tensor(dx), Tensor(dw), Tensor(db)
    if emit_cl and _cl_ok(x):
        dx_cl = DeviceArray.empty((L.cpt_channels_last_bytes(N, C, HW, 1, _lib.MODE_BF16),), np.uint8)
        csum = DeviceArray.empty((C,), np.float32) if emit_sum else None
        ws, wsb = workspace(L.cpt_bn_cl_workspace_size(N, C, HW))
        _lib.check(L.cpt_bn_act_bwd_cl(f32ptr(x), f32ptr(dy), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, dx.ptr, dx_cl.ptr,
                                       csum.ptr if csum is not None else None, dw.ptr, db.ptr, N, C, HW, act, ws, wsb, stream_ptr()))
        dx.cl = (_lib.MODE_BF16, dx_cl, csum)
        return Tensor(dx), Tensor(dw), Tensor(db)
    ws, wsb = workspace(L.cpt_bn_workspace_size(N, C, HW))
    _lib.check(L.cpt_bn_act_bwd(f32ptr(x), f32ptr(dy), f32ptr(w), f32ptr(b), save_mean.ptr, save_rstd.ptr, dx.ptr, dw.ptr, db.ptr,
                                N, C, HW, act, ws, wsb, stream_ptr()))
    return Tensor(dx), Tensor(dw), Tensor(db)


class BatchNorm2DFn(Function):
    """(:120-177) statistics over (0, 2, 3)."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, rmean: Tensor, rvar: Tensor, w: Tensor, b: Tensor, m: float, eps: float,
                training: bool, emit_cl: bool = False, skip: Tensor = None, relu_cache=None,
                pool2: bool = False) -> tuple[Tensor, Tensor, Tensor]:
        """``emit_cl`` (extension): also write y as channels-last bf16 for a tensor-core convolution that consumes it.
        ``skip`` (extension): return ``relu(bn(x) + skip)`` instead (residual tail, see ``_bn_forward``).
        ``pool2`` (extension): return ``maxpool2(relu(bn(x)))`` instead; ``backward`` then takes the pooled gradient."""
        if x.ndim != 4:
            raise ShapeError(f"Expected input to be 4D, got {x.ndim}D.")
        B, C, H, W = x.shape
        return _bn_forward(cache, x, rmean, rvar, w, b, m, eps, training, B, C, H * W, ACT_NONE, emit_cl and skip is None, skip,
                           relu_cache, pool2)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, dw_out=None, db_out=None, emit_cl: bool = False,
                 emit_sum: bool = False) -> tuple[Tensor, Tensor, Tensor]:
        return _bn_backward(cache, dy, dw_out, db_out, emit_cl, emit_sum)


class BatchNorm1DFn(Function):
    """(:10-70) statistics over (0,) for 2-D input, (0, 2) for 3-D input."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, rmean: Tensor, rvar: Tensor, w: Tensor, b: Tensor, m: float, eps: float,
                training: bool) -> tuple[Tensor, Tensor, Tensor]:
        if x.ndim not in {2, 3}:
            raise ShapeError(f"Expected input to be 2D or 3D, got {x.ndim}D.")
        N, C = x.shape[0], x.shape[1]
        HW = x.shape[2] if x.ndim == 3 else 1
        return _bn_forward(cache, x, rmean, rvar, w, b, m, eps, training, N, C, HW)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, dw_out=None, db_out=None) -> tuple[Tensor, Tensor, Tensor]:
        return _bn_backward(cache, dy, dw_out, db_out)


class BatchNormReLU2DFn(Function):
    """``ReLUFn.forward(BatchNorm2DFn.forward(x))`` in one pass (no reference counterpart: this is what the Sequential
    peephole BatchNorm2D -> ReLU calls).  Same signatures as BatchNorm2DFn; ``backward`` takes the gradient w.r.t. the
    ReLU OUTPUT and folds ``dy * (y > 0)`` (activation_funcs.py:32-34) into both batch-norm backward passes."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, rmean: Tensor, rvar: Tensor, w: Tensor, b: Tensor, m: float, eps: float,
                training: bool, emit_cl: bool = False) -> tuple[Tensor, Tensor, Tensor]:
        if x.ndim != 4:
            raise ShapeError(f"Expected input to be 4D, got {x.ndim}D.")
        B, C, H, W = x.shape
        return _bn_forward(cache, x, rmean, rvar, w, b, m, eps, training, B, C, H * W, ACT_RELU, emit_cl)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, dw_out=None, db_out=None, emit_cl: bool = False,
                 emit_sum: bool = False) -> tuple[Tensor, Tensor, Tensor]:
        return _bn_backward(cache, dy, dw_out, db_out, emit_cl, emit_sum)


class BatchNormReLU1DFn(Function):
    """1-D counterpart of BatchNormReLU2DFn (statistics over (0,) / (0, 2))."""

    @staticmethod
    def forward(cache: FunctionCache, x: Tensor, rmean: Tensor, rvar: Tensor, w: Tensor, b: Tensor, m: float, eps: float,
                training: bool) -> tuple[Tensor, Tensor, Tensor]:
        if x.ndim not in {2, 3}:
            raise ShapeError(f"Expected input to be 2D or 3D, got {x.ndim}D.")
        N, C = x.shape[0], x.shape[1]
        HW = x.shape[2] if x.ndim == 3 else 1
        return _bn_forward(cache, x, rmean, rvar, w, b, m, eps, training, N, C, HW, ACT_RELU)

    @staticmethod
    def backward(cache: FunctionCache, dy: Tensor, dw_out=None, db_out=None) -> tuple[Tensor, Tensor, Tensor]:
        return _bn_backward(cache, dy, dw_out, db_out)


def batchnorm1d(x, rmean, rvar, w, b, m: float = 0.1, eps: float = 1e-5, training: bool = False):
    return BatchNorm1DFn.forward(PseudoCache(), x, rmean, rvar, w, b, m, eps, training)


def batchnorm2d(x, rmean, rvar, w, b, m: float = 0.1, eps: float = 1e-5, training: bool = False):
    return BatchNorm2DFn.forward(PseudoCache(), x, rmean, rvar, w, b, m, eps, training)
