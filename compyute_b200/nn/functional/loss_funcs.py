"""Cross-entropy from logits.  API of compyute/nn/functional/loss_funcs.py:53-95 (+ accuracy, metric_funcs.py:10-25)."""

from __future__ import annotations

import numpy as np

from ... import _lib
from ...tensors import DeviceArray, ShapeError, Tensor, f32ptr, require_cuda, stream_ptr
from .functions import Function, FunctionCache, PseudoCache

__all__ = ["cross_entropy_loss", "CrossEntropyLossFn", "accuracy_score"]


def _targets_i32(targets: Tensor) -> DeviceArray:
    require_cuda(targets)
    if targets.data.dtype == np.int32:
        return targets.data
    if targets.data.dtype == np.int64:  # labels arrive as int64 like in the reference; the kernels index with int32
        from ... import device_ops as D
        return D.astype(targets.data, np.int32)  # cast kernel on the device: no host round trip
    raise ValueError(f"Input must be an integer, got '{targets.data.dtype}'.")


class CrossEntropyLossFn(Function):
    """softmax → -mean log(p_target + eta) (:57-64); backward (p - onehot) / B (:67-69).  The loss stays on the
    device as a 1-element tensor; ``.item()`` is the only host sync."""

    @staticmethod
    def forward(cache: FunctionCache, logits: Tensor, targets: Tensor, eta: float) -> Tensor:
        require_cuda(logits)
        if logits.ndim < 2:
            raise ShapeError(f"Expected logits to be at least 2D (..., classes), got {logits.ndim}D.")
        # (..., classes): softmax over the last dim, mean over all leading dims (loss_funcs.py:57-64) == the 2-D case on
        # the flattened rows; backward divides by the number of rows, math.prod(shape[:-1]) (:67-69)
        NC = logits.shape[-1]
        B = logits.size // NC
        t32 = _targets_i32(targets)
        if t32.size != B:
            raise ShapeError(f"Expected {B} targets for logits of shape {logits.shape}, got {targets.shape}.")
        probs = DeviceArray.empty((B, NC), np.float32)
        loss = DeviceArray.empty((1,), np.float32)
        rows = DeviceArray.empty((B,), np.float32)
        _lib.check(_lib.lib().cpt_softmax_ce_fwd(f32ptr(logits), t32.ptr, probs.ptr, loss.ptr, rows.ptr, B, NC, float(eta),
                                                 stream_ptr()))
        cache.push(t32, probs, logits.shape)
        return Tensor(loss.reshape(()))

    @staticmethod
    def backward(cache: FunctionCache) -> Tensor:
        t32, probs, shape = cache.pop()
        B, NC = probs.shape
        d = DeviceArray.empty((B, NC), np.float32)
        _lib.check(_lib.lib().cpt_softmax_ce_bwd(probs.ptr, t32.ptr, d.ptr, B, NC, stream_ptr()))
        return Tensor(d.reshape(shape))


def cross_entropy_loss(logits: Tensor, targets: Tensor, eta: float = 1e-8) -> Tensor:
    return CrossEntropyLossFn.forward(PseudoCache(), logits, targets, eta)


def accuracy_score(logits: Tensor, targets: Tensor) -> float:
    """mean(argmax(logits, -1) == targets) (metric_funcs.py:10-25)."""
    require_cuda(logits)
    NC = logits.shape[-1]
    B = logits.size // NC
    t32 = _targets_i32(targets)
    cnt = DeviceArray.empty((1,), np.int32)
    _lib.check(_lib.lib().cpt_accuracy_count(f32ptr(logits), t32.ptr, cnt.ptr, B, NC, stream_ptr()))
    return cnt.item() / B
