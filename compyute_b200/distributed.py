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
           "shard_batch", "shard_bounds", "barrier", "set_sync_batchnorm", "sync_batchnorm_active", "SymmetricArena",
           "symmetric_memory_available", "use_own_nccl", "plan_shards", "bind_to_gpu_numa_node", "shard_weight"]

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
    if backend == "nccl" and d.get_world_size() > 1 and os.environ.get("CPT_OWN_NCCL", "1") != "0":
        try:  # gradient all-reduces go through the library's own communicator (C ABI cpt_nccl_*); the group stays the side channel
            use_own_nccl(True)
        except Exception as e:  # pragma: no cover - depends on the NCCL build
            import warnings
            warnings.warn(f"compyute_b200: own NCCL communicator unavailable ({e}); using torch.distributed for all-reduces")


def bind_to_gpu_numa_node(device_index: int | None = None) -> dict:
    """Pins this process (CPU affinity + preferred memory node) to the NUMA node its GPU hangs off.  With one process per GPU
    the pinned staging buffers of all ranks otherwise land wherever the launcher happened to run: half of the host<->device
    traffic of an 8-GPU box then crosses the socket interconnect, which is what bounds the end-to-end (host-buffer) numbers
    at 8 ranks.  Call BEFORE allocating pinned memory.  Returns what was done (for the benchmark record); never raises."""
    info: dict = {"bound": False}
    try:
        import ctypes
        import torch
        idx = torch.cuda.current_device() if device_index is None else device_index
        pr = torch.cuda.get_device_properties(idx)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info.update(pci=bdf, node=node)
        if node < 0:
            return info
        cpus: set[int] = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
        mask = ctypes.c_ulong(1 << node)  # set_mempolicy(MPOL_PREFERRED = 1, &mask, maxnode)
        rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), 64)
        info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
        info["bound"] = bool(allowed)
    except Exception as e:  # pragma: no cover - best effort
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def is_initialized() -> bool:
    return _dist is not None and _dist.is_initialized()


def rank() -> int:
    return _d().get_rank() if is_initialized() else 0


def world_size() -> int:
    return _d().get_world_size() if is_initialized() else 1


def barrier() -> None:
    if is_initialized():
        _d().barrier()


_own_nccl = False


def use_own_nccl(enable: bool = True) -> bool:
    """Routes ``all_reduce_sum`` of fp32 device arrays through the library's own NCCL communicator (C ABI ``cpt_nccl_*``,
    csrc/nccl_comm.cu) instead of ``torch.distributed``: rank 0 creates the unique id, the process group — used as the side
    channel only — broadcasts its 128 bytes, every rank calls ``cpt_nccl_init``.  Returns whether the communicator is up."""
    global _own_nccl
    from . import _lib
    L = _lib.lib()
    if not enable:
        if _own_nccl:
            _lib.check(L.cpt_nccl_destroy())
        _own_nccl = False
        return False
    if _own_nccl or not is_initialized() or world_size() == 1:
        return _own_nccl
    import ctypes
    import torch
    uid = (ctypes.c_ubyte * 128)()
    if rank() == 0:
        _lib.check(L.cpt_nccl_unique_id(uid))
    t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda" if _d().get_backend() == "nccl" else "cpu")
    _d().broadcast(t, src=0)
    uid = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
    _lib.check(L.cpt_nccl_init(rank(), world_size(), uid))
    _own_nccl = True
    return True


def all_reduce_sum(arr) -> None:
    """In-place SUM all-reduce of a DeviceArray (or a torch tensor) across ranks."""
    if not is_initialized() or world_size() == 1:
        return
    buf = getattr(arr, "_buf", arr)
    if _own_nccl and str(buf.dtype) == "torch.float32" and buf.is_cuda and buf.is_contiguous():
        import torch
        from . import _lib
        _lib.check(_lib.lib().cpt_nccl_allreduce_sum_f32(buf.data_ptr(), buf.numel(), torch.cuda.current_stream().cuda_stream))
        return
    _d().all_reduce(buf, op=_d().ReduceOp.SUM)


def all_gather(arr):
    """Gathers a DeviceArray from every rank into a new DeviceArray of shape (world, *arr.shape), in rank order."""
    from .tensors import DeviceArray
    out = DeviceArray.empty((world_size(),) + tuple(arr.shape), arr.dtype)
    _d().all_gather_into_tensor(out._buf.view(-1), arr._buf.view(-1))
    return out


def plan_shards(total_elems: int, world: int, align: int = 64) -> tuple[int, int]:
    """Layout of a flat arena that ``world`` ranks update in equal contiguous shards: returns (padded total, shard size),
    both multiples of ``align`` elements (``align`` >= 4: the fused step moves 16-byte vectors)."""
    shard = (total_elems + world * align - 1) // (world * align) * align
    return shard * world, shard


_symm_state = {"checked": False, "ok": False, "why": ""}


def symmetric_memory_available() -> bool:
    """True when torch's symmetric-memory allocator can map one buffer into every rank of the (NCCL) process group — the
    plumbing under the fused data-parallel optimizer step (csrc/dp_step.cu).  ``CPT_DP_FUSED_STEP=0`` switches it off."""
    if _symm_state["checked"]:
        return _symm_state["ok"]
    _symm_state["checked"] = True
    if os.environ.get("CPT_DP_FUSED_STEP", "1") == "0":
        _symm_state["why"] = "disabled by CPT_DP_FUSED_STEP=0"
        return False
    if not is_initialized() or world_size() == 1 or _d().get_backend() != "nccl":
        _symm_state["why"] = "needs an initialised NCCL process group with more than one rank"
        return False
    try:
        probe = SymmetricArena(1024)
        _symm_state["ok"] = True
        _symm_state["why"] = "multicast (NVLS)" if probe.multicast_ptr else "peer mappings only (no multicast support)"
        del probe
    except Exception as e:  # pragma: no cover - depends on the driver / fabric
        _symm_state["why"] = f"{type(e).__name__}: {e}"
    return _symm_state["ok"]


def symmetric_memory_note() -> str:
    return _symm_state["why"]


class SymmetricArena:
    """A flat fp32 buffer of ``n_elems`` elements in SYMMETRIC memory: allocated with the same size on every rank
    (collective call), mapped into every peer's address space and — where the NVSwitch fabric supports it — into one
    multicast object (``multicast_ptr``; 0 otherwise).  Device memory, the mappings and the stream-ordered barrier come from
    ``torch.distributed._symmetric_memory`` (plumbing); the kernels that read / write through them are the library's."""

    def __init__(self, n_elems: int) -> None:
        import numpy as np
        import torch
        import torch.distributed._symmetric_memory as symm
        from .tensors import DeviceArray
        dev = torch.device("cuda", torch.cuda.current_device())
        self.tensor = symm.empty(int(n_elems), dtype=torch.float32, device=dev)
        self.handle = symm.rendezvous(self.tensor, group=_d().group.WORLD)
        self.tensor.zero_()
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.peer_ptrs_dev = int(self.handle.buffer_ptrs_dev)   # device array of world base addresses
        self.array = DeviceArray(self.tensor, (int(n_elems),), np.float32)

    def barrier(self, channel: int = 0) -> None:
        """Cross-rank barrier ordered on the current stream (a tiny kernel that signals every peer and waits for them)."""
        self.handle.barrier(channel=channel)


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


def shard_weight(n_local: int, n_global: int, world: int | None = None) -> float:
    """Weight of this rank's (local-mean) gradient in the global-batch mean when shards are uneven: n_local * world / n_global.
    The exchange averages the ranks' gradients with equal weights (sum, then 1/world); a rank that holds more samples than the
    others must count for more — set ``optimizer.local_weight = shard_weight(...)`` (1.0 for equal shards)."""
    world = world_size() if world is None else world
    return float(n_local) * world / float(n_global)


def shard_batch(*arrays: np.ndarray):
    """Slices host arrays along dim 0 to this rank's shard (the natural shard point is the Dataloader, dataloaders.py:62-66)."""
    out = []
    for a in arrays:
        lo, hi = shard_bounds(a.shape[0])
        out.append(a[lo:hi])
    return out[0] if len(out) == 1 else tuple(out)
