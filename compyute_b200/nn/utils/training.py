"""Training utilities (compyute/nn/utils/training.py)."""

from __future__ import annotations

import numpy as np

from ...tensors import DeviceArray

__all__ = ["clip_grad_norm"]


def clip_grad_norm(parameters, max_norm: float) -> float:
    """compyute/nn/utils/training.py:12-39: scales all gradients so that their joint L2 norm is at most ``max_norm``;
    returns the unclipped norm.  In a data-parallel run the gradients are exchanged and averaged first (``Optimizer.sync_grads``),
    so the result equals clipping the global-batch gradient of a single process.  On the device: one sum-of-squares reduction per gradient, one 4-byte D2H for the norm (the
    reference concatenates every gradient into one array first), one in-place scale per gradient when clipping."""
    params = [p for p in parameters if p.grad]
    if not params:
        return 0.0
    from ... import device_ops as D
    from ..optimizers import slot_owner
    # data-parallel runs: the norm is that of the world-mean gradient (every rank then applies the SAME coefficient).  The
    # owning optimizers finish their exchange first (this also waits for overlapped bucket all-reduces still in flight
    # before the in-place scaling below touches the arena) and will not exchange again in step().
    for opt in {id(o): o for o in (slot_owner(p) for p in params) if o is not None}.values():
        opt.sync_grads()
    sq = 0.0
    parts = []
    for p in params:
        g = p.grad.data
        if isinstance(g, DeviceArray):
            parts.append(D.reduce("sumsq", g))
        else:
            sq += float(np.sum(np.square(g, dtype=np.float64)))
    if parts:
        sq += float(D.reduce("sum", D.concat([s.reshape(1) for s in parts], 0)).item())
    grad_norm = float(np.sqrt(sq))
    if grad_norm <= max_norm:
        return grad_norm
    clip_coef = max_norm / grad_norm
    for p in params:
        p.grad *= clip_coef
    return grad_norm
