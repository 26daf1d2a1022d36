"""Batch-sharded data parallelism: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch) as the
plumbing.  The reference has no distributed code at all; the path shards on the batch dimension and has
exactly one exchange step per train step — a SUM all-reduce of the flat gradient arena — which
``Optimizer.step`` issues before the fused update.

Rank r of n takes samples [r*B/n, (r+1)*B/n) of a global minibatch (``shard_batch``); parameters are made identical
at start (same seed, or ``broadcast_parameters``).  BatchNorm statistics are per shard (DDP semantics): the parity
oracle for DP is "reference run per shard, gradients averaged".
"""

from __future__ import annotations

import os
from typing import Iterable

import numpy as np

__all__ = ["init", "is_initialized", "rank", "world_size", "all_reduce_sum", "all_reduce_sum_async", "all_gather", "broadcast_parameters",
           "shard_batch", "shard_bounds", "barrier", "set_sync_batchnorm", "sync_batchnorm_active"]

_dist = None


def _d():
    global _dist
    if _dist is None:
        import torch.distributed as dist
        _dist = dist
    return _dist


def init(backend: str | None = None) -> None:
    """Initialises the default process group from torchrun's env (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    import torch
    d = _d()
    if d.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    d.init_process_group(backend=backend, rank=int(os.environ.get("RANK", "0")),
                         world_size=int(os.environ.get("WORLD_SIZE", "1")))


def is_initialized() -> bool:
    return _dist is not None and _dist.is_initialized()


def rank() -> int:
    return _d().get_rank() if is_initialized() else 0


def world_size() -> int:
    return _d().get_world_size() if is_initialized() else 1


def barrier() -> None:
    if is_initialized():
        _d().barrier()


def all_reduce_sum(arr) -> None:
    """In-place SUM all-reduce of a DeviceArray (or a torch tensor) across ranks."""
    if not is_initialized() or world_size() == 1:
        return
    buf = getattr(arr, "_buf", arr)
    _d().all_reduce(buf, op=_d().ReduceOp.SUM)


def all_gather(arr):
    """Gathers a DeviceArray from every rank into a new DeviceArray of shape (world, *arr.shape), in rank order."""
    from .tensors import DeviceArray
    out = DeviceArray.empty((world_size(),) + tuple(arr.shape), arr.dtype)
    _d().all_gather_into_tensor(out._buf.view(-1), arr._buf.view(-1))
    return out


_sync_bn = False


def set_sync_batchnorm(enabled: bool) -> None:
    """Synchronised BatchNorm: in training mode every BatchNorm1D/2D takes its batch statistics — and the two
    sums of its backward pass — over the GLOBAL batch of all ranks instead of the local shard, which makes data-parallel
    training at n ranks equal to single-process training on the full batch.  Costs one small all-gather (3·C floats per
    rank) per forward and one all-reduce (2·C floats) per backward and layer; off by default (per-shard statistics, DDP
    semantics).  No effect without an initialised process group."""
    global _sync_bn
    _sync_bn = bool(enabled)


def sync_batchnorm_active() -> bool:
    return _sync_bn and is_initialized() and world_size() > 1


def all_reduce_sum_async(buf):
    """Asynchronous SUM all-reduce of a torch tensor view: NCCL's stream first waits for the work already enqueued on
    the current stream, then runs next to whatever is enqueued afterwards.  Returns the work handle (``.wait()`` makes
    the current stream wait for it)."""
    return _d().all_reduce(buf, op=_d().ReduceOp.SUM, async_op=True)


def broadcast_parameters(tensors: Iterable, src: int = 0) -> None:
    """Makes every rank start from rank ``src``'s parameters / buffers."""
    if not is_initialized() or world_size() == 1:
        return
    for t in tensors:
        _d().broadcast(t.data._buf, src=src)


def shard_bounds(n: int, r: int | None = None, world: int | None = None) -> tuple[int, int]:
    """[lo, hi) of rank ``r``'s contiguous shard of ``n`` samples (remainder spread over the first ranks)."""
    r = rank() if r is None else r
    world = world_size() if world is None else world
    base, rem = divmod(n, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def shard_batch(*arrays: np.ndarray):
    """Slices host arrays along dim 0 to this rank's shard (the natural shard point is the Dataloader, dataloaders.py:62-66)."""
    out = []
    for a in arrays:
        lo, hi = shard_bounds(a.shape[0])
        out.append(a[lo:hi])
    return out[0] if len(out) == 1 else tuple(out)
