"""Optimizers with a fused multi-tensor step and batch-sharded data parallelism.

API of compyute/nn/optimizers.py:14-92 (``Optimizer``), :95-176 (``SGD``), :179-271 (``Adam``), :274-362 (``AdamW``):
same constructor arguments, ``step() / reset_grads() / get_state_dict() / load_state_dict()``, ``t`` starting at 1,
state layout ``{i: {"m": Tensor, "v": Tensor}}``.  What changes is *how* a step runs:

* one kernel launch updates every parameter (``cpt_adam_step`` / ``cpt_sgd_step`` over a device pointer table)
  instead of ~12 launches and ~10 temporaries per parameter;
* gradients live in one flat fp32 arena (``Parameter.grad_slot`` views; layers write dW/db straight into it), so
  in data-parallel mode ``step()`` issues ONE sum all-reduce (NCCL over NVLink) on the arena and folds the 1/world
  averaging into the update kernel's ``grad_scale`` (local CE gradients are already means over the local shard,
  loss_funcs.py:69).

``lr`` and ``t`` stay plain Python attributes read at every step, so LR schedulers keep working (Appendix A.17).
"""

from __future__ import annotations

import ctypes
import os
import weakref
from typing import Any, Iterable, Optional

import numpy as np

from .. import _lib, distributed, graph
from ..tensors import DeviceArray, Tensor, stream_ptr
from .parameter import Parameter

__all__ = ["Optimizer", "SGD", "Adam", "AdamW", "NAdam", "plan_grad_buckets"]


_SLOT_OWNERS: dict[int, tuple[Any, int]] = {}  # arena slot pointer -> (weakref to its optimizer, parameter index)


def notify_grad_written(parameter: Parameter, first: bool) -> None:
    """Hook for ``Module.update_parameter_grad``: tells the owning optimizer that ``parameter``'s arena slot holds this
    step's gradient (drives the overlapped data-parallel exchange; a no-op otherwise)."""
    slot = getattr(parameter, "grad_slot", None)
    if slot is None or parameter.grad is None or getattr(parameter.grad.data, "ptr", None) != slot.ptr:
        return
    owner = _SLOT_OWNERS.get(slot.ptr)
    if owner is None:
        return
    opt = owner[0]()
    if opt is not None and opt.overlap_grad_sync:
        opt._on_grad_ready(owner[1], first)


def slot_owner(parameter: Parameter):
    """The optimizer whose gradient arena holds ``parameter``'s slot, or None."""
    slot = getattr(parameter, "grad_slot", None)
    owner = _SLOT_OWNERS.get(slot.ptr) if slot is not None else None
    return owner[0]() if owner is not None else None


def plan_grad_buckets(offsets: list[int], sizes: list[int], bucket_elems: int) -> list[tuple[int, int, list[int]]]:
    """Contiguous buckets over the gradient arena for the overlapped data-parallel exchange.  Backward produces gradients
    from the LAST parameter to the first, so buckets are cut walking the arena from its end: returns
    ``[(lo, hi, [param indices]), ...]`` in expected order of completion, each covering >= ``bucket_elems`` elements
    (except possibly the last one, at the front of the arena); together they tile ``[0, arena size)`` exactly."""
    buckets: list[tuple[int, int, list[int]]] = []
    hi = (offsets[-1] + (sizes[-1] + 63) // 64 * 64) if offsets else 0
    members: list[int] = []
    for i in range(len(offsets) - 1, -1, -1):
        members.append(i)
        if hi - offsets[i] >= bucket_elems or i == 0:
            buckets.append((offsets[i], hi, members))
            hi, members = offsets[i], []
    return buckets


class Optimizer:
    """Optimizer base class (optimizers.py:14-92)."""

    _state_keys: tuple[str, ...] = ()
    _fused_dp_keys: Optional[tuple[str, ...]] = None  # state buffers of the fused data-parallel step (None: not supported)

    def __init__(self, parameters: Optional[Iterable[Parameter]] = None, lr: float = 1e-3) -> None:
        self.lr = lr
        self.t = 1
        self._parameters: list[Parameter] = []
        self._state: dict[int, dict[str, Tensor]] = {}
        self._arena: Optional[DeviceArray] = None
        self._table_dev: Optional[DeviceArray] = None
        self._table_key: Optional[bytes] = None
        self._data_parallel = True  # all-reduce gradients in step() whenever a process group with world > 1 exists
        # overlapped exchange: buckets of the arena are all-reduced (async, NCCL's own stream) as soon as backward has
        # produced every gradient in them, instead of one all-reduce at step(); opt-in, needs one gradient per parameter
        # and step (no shared parameters)
        self.overlap_grad_sync = False
        self.bucket_bytes = 32 << 20
        self.reserve_sms = 16       # SMs the persistent tensor-core kernels leave to NCCL while all-reduces are in flight
        self._buckets = None
        self._bucket_pending: list[int] = []
        self._bucket_work: list[Any] = []
        self._bucket_launched: list[bool] = []
        self._live = None           # device float[8]: per-step scalars read by the update kernel in CUDA-graph replays
        self._synced = False        # sync_grads() already exchanged (and averaged) this step's gradients
        # fused data-parallel step (csrc/dp_step.cu): gradient + parameter arenas in symmetric memory, each rank updates its
        # shard from the in-switch gradient sum and multicasts the new parameters; None = classic all-reduce + replicated update
        self.local_weight = 1.0     # uneven data-parallel shards: distributed.shard_weight(n_local, n_global); 1.0 = equal shards
        self.fused_dp_step = True   # use it whenever the process group supports it (set False before the first step to opt out)
        self._fused, self._fused_checked = None, False
        if parameters is not None:
            self.set_parameters(parameters)

    # ---- reference API -------------------------------------------------------------------------
    def set_parameters(self, parameters: Iterable[Parameter]) -> None:
        """De-duplicates by ``p.ptr`` (optimizers.py:36-53) and lays out the flat gradient arena."""
        seen: set[int] = set()
        self._parameters = []
        for p in parameters:
            if p.ptr in seen:
                continue
            self._parameters.append(p)
            seen.add(p.ptr)
        self._state = {i: {} for i in range(len(self._parameters))}
        self._build_arena()

    def get_state_dict(self) -> dict[str, dict[Any, Any]]:
        if self._fused is not None:
            self._materialize_fused_state()
        skip = {"_parameters", "_state", "_arena", "_table_dev", "_table_key", "_data_parallel", "_live", "_synced", "_fused", "_fused_checked", "fused_dp_step", "local_weight", "overlap_grad_sync",
                "bucket_bytes", "reserve_sms", "_buckets", "_bucket_pending", "_bucket_work", "_bucket_launched", "_offsets", "_bucket_of"}
        return {"state": self._state, "vars": {k: v for k, v in vars(self).items() if k not in skip}}

    def load_state_dict(self, state_dict: dict[str, dict[Any, Any]]) -> None:
        self._state = state_dict["state"]
        for k, v in state_dict["vars"].items():
            setattr(self, k, v)
        self._table_key = None
        if self._fused is not None:
            self._scatter_fused_state()

    def reset_grads(self) -> None:
        """``p.grad = None`` (optimizers.py:85-88); arena slots are simply overwritten by the next backward."""
        for p in self._parameters:
            p.grad = None
        self._reset_buckets()
        self._fused_reset_buckets()
        self._synced = False

    def sync_grads(self) -> None:
        """Data-parallel runs: completes the gradient exchange NOW and leaves the world-MEAN gradient in every ``p.grad``
        (``step()`` then skips the exchange).  Anything that reads or rescales gradients before ``step()`` — gradient
        clipping, logging of gradient norms — must see the averaged global gradient, not this rank's shard gradient:
        ``clip_grad_norm`` calls this itself.  No-op on one rank."""
        if self._dp_world() == 1 or self._synced or self._arena is None:
            return
        if self._fused is not None:  # plain all-reduce of the symmetric gradient arena; the fused step then skips its own sum
            self._gather_grads_into_arena()
            if self.local_weight != 1.0:
                self._arena *= float(self.local_weight)
            distributed.all_reduce_sum(self._arena)
            scale = 1.0 / self._dp_world()
        else:
            scale = self._sync_grads()
        self._arena *= float(scale)
        self._synced = True

    def step(self) -> None:
        """Updates the parameters (one fused launch).  Inside a CUDA-graph capture the per-step scalars are read from device
        memory and ``t`` is advanced by ``upload_live_scalars`` at replay time instead."""
        if not self._fused_checked and not graph.is_capturing():
            self._maybe_enable_fused()
        if self._fused is not None and self._dp_world() > 1:
            if graph.is_capturing():
                # captured data-parallel step: barriers and the fused kernel are ordinary launches on the capture stream; the
                # per-step scalars come from the live buffer and ``t`` advances in upload_live_scalars at replay time
                self._fused_step(self._peek_scalars(), self._live_buffer().ptr)
                return
            self._fused_step(None, None)
            self._synced = False
            self.t += 1
            return
        scale = 1.0 if self._synced else self._sync_grads()
        self._synced = False
        if graph.is_capturing():
            self._launch(self._peek_scalars(), self._live_buffer().ptr, scale)
            return
        self._launch(self._step_scalars(), None, scale)
        self.t += 1

    def _step_scalars(self) -> list[float]:
        """Per-step scalars [lr, ...] for the current ``t``; may advance internal products (NAdam)."""
        raise NotImplementedError

    def _peek_scalars(self) -> list[float]:
        """Like ``_step_scalars`` but without side effects (placeholder values baked into a captured launch)."""
        return [float(self.lr)] + [1.0] * 5

    def _launch(self, sc: list[float], live, scale: float) -> None:
        raise NotImplementedError

    def _live_buffer(self) -> DeviceArray:
        """Must exist BEFORE a capture starts (an allocation + memset inside the capture would be replayed every time)."""
        if self._live is None:
            if graph.is_capturing():
                raise RuntimeError("optimizer live-scalar buffer must be created before capture (use cp.graph.CapturedStep)")
            self._live = DeviceArray.zeros((8,), np.float32)
        return self._live

    def upload_live_scalars(self) -> None:
        """Called before each replay of a captured step: scalars of step ``t`` -> device buffer, then ``t += 1``."""
        import ctypes
        sc = self._step_scalars()
        vals = (ctypes.c_float * len(sc))(*sc)
        _lib.check(_lib.lib().cpt_set_live_scalars(self._live_buffer().ptr, vals, len(sc), stream_ptr()))
        self.t += 1

    # ---- arena / pointer table -----------------------------------------------------------------
    def _build_arena(self) -> None:
        """One fp32 buffer holding every gradient, 64-element aligned slots.  Only for cuda parameters."""
        self._arena = None
        cuda_params = [p for p in self._parameters if isinstance(p.data, DeviceArray)]
        if not cuda_params or len(cuda_params) != len(self._parameters):
            return
        offs, total = [], 0
        for p in cuda_params:
            offs.append(total)
            total += (p.size + 63) // 64 * 64
        self._fused, self._fused_checked = None, False  # the symmetric arenas are built at the first data-parallel step()
        self._arena = DeviceArray.zeros((total,), np.float32)
        self._offsets = offs
        for ptr in [k for k, (ref, _) in _SLOT_OWNERS.items() if ref() is None or ref() is self]:
            del _SLOT_OWNERS[ptr]  # slots of collected optimizers / of this optimizer's previous arena
        flat = self._arena._buf
        for i, (p, o) in enumerate(zip(cuda_params, offs)):
            p.grad_slot = DeviceArray(flat[o:o + p.size], p.shape, np.float32)
            _SLOT_OWNERS[p.grad_slot.ptr] = (weakref.ref(self), i)  # looked up by Module.update_parameter_grad
        self._buckets = None

    # ---- fused data-parallel step ---------------------------------------------------------------
    def _maybe_enable_fused(self) -> None:
        """First data-parallel step(): move the gradient arena G and the parameters (arena P) into symmetric memory.  Collective —
        every rank of the group reaches its first step() —, which is why it does not happen in set_parameters (an optimizer
        built on one rank only, e.g. a single-process reference run next to a DP run, must not start a rendezvous).

        The arenas are cut into BUCKETS of whole parameters (one bucket without ``overlap_grad_sync``; ``plan_grad_buckets`` of
        >= ``bucket_bytes`` each with it, in the order backward completes them), every bucket padded to a multiple of
        world * 64 elements so that rank r owns the r-th equal slice of EVERY bucket and keeps the moments of those slices only.
        ``p.data`` / ``p.grad_slot`` become views of P / G (the fused step writes every replica's P through the multicast /
        peer mappings)."""
        self._fused_checked = True
        if not (self.fused_dp_step and self._fused_dp_keys is not None and self._arena is not None and self._dp_world() > 1
                and distributed.symmetric_memory_available()):
            return
        if any(self._state[i] for i in self._state):
            return  # moments already exist in the replicated layout (resumed run): keep the classic path
        world, rank = distributed.world_size(), distributed.rank()
        sizes = [p.size for p in self._parameters]
        if self.overlap_grad_sync:
            plan = plan_grad_buckets(self._offsets, sizes, max(1, self.bucket_bytes // 4))
        else:
            plan = [(0, self._arena.size, list(range(len(sizes) - 1, -1, -1)))]
        # new layout: buckets in arena order, parameters 64-aligned inside, bucket length a multiple of world * 64
        new_off, buckets, cursor = [0] * len(sizes), [], 0
        for lo, hi, members in sorted(plan, key=lambda b: b[0]):
            b_lo = cursor
            for i in sorted(members):
                new_off[i] = cursor
                cursor += (sizes[i] + 63) // 64 * 64
            length, shard = distributed.plan_shards(cursor - b_lo, world)
            cursor = b_lo + length
            buckets.append({"lo": b_lo, "hi": cursor, "shard": shard, "off": b_lo + rank * shard, "members": sorted(members),
                            "state": {}, "pending": len(members), "launched": False, "order": plan.index((lo, hi, members))})
        buckets.sort(key=lambda b: b["order"])  # expected completion order of backward
        G, P = distributed.SymmetricArena(cursor), distributed.SymmetricArena(cursor)
        old = self._arena._buf
        for ptr in [k for k, (ref, _) in _SLOT_OWNERS.items() if ref() is self]:
            del _SLOT_OWNERS[ptr]
        for i, (p, o_old, o) in enumerate(zip(self._parameters, self._offsets, new_off)):
            G.tensor[o:o + p.size].copy_(old[o_old:o_old + p.size])  # this step's gradients are already in the old arena
            P.tensor[o:o + p.size].copy_(p.data._buf.reshape(-1))
            p.data = DeviceArray(P.tensor[o:o + p.size], p.shape, np.float32)
            had_slot_grad = p.grad is not None and p.grad_slot is not None and p.grad.data.ptr == p.grad_slot.ptr
            p.grad_slot = DeviceArray(G.tensor[o:o + p.size], p.shape, np.float32)
            _SLOT_OWNERS[p.grad_slot.ptr] = (weakref.ref(self), i)
            if had_slot_grad:
                p.grad = Tensor(p.grad_slot)
        self._arena, self._offsets = G.array, new_off
        self._table_key = None
        self._buckets = None
        # multimem pulls EVERY member's copy through the switch, the requester's own included: at 2 ranks that is 1.5x the bytes
        # of plain peer loads / stores (measured, 537 MB: 1.43 vs 1.06 ms), from 3 ranks on it is the cheaper path
        env = os.environ.get("CPT_DP_MULTIMEM")
        use_mc = bool(P.multicast_ptr and G.multicast_ptr) and (world > 2 if env is None else env != "0")
        self._fused = {"G": G, "P": P, "world": world, "buckets": buckets, "use_multimem": use_mc, "side": None, "scalars": None,
                       "bucket_of": {i: b for b in buckets for i in b["members"]}}

    def _fused_view(self, bucket: dict, small_grid: bool) -> "_lib.DpView":
        f = self._fused
        v = _lib.DpView()
        v.p_local, v.g_local = f["P"].array.ptr, f["G"].array.ptr
        mc = f["use_multimem"]
        v.p_mc, v.g_mc = (f["P"].multicast_ptr if mc else None), (f["G"].multicast_ptr if mc else None)
        v.p_peers, v.g_peers = f["P"].peer_ptrs_dev, f["G"].peer_ptrs_dev
        v.shard_off, v.shard_elems, v.world = bucket["off"], bucket["shard"], f["world"]
        v.pre_reduced = 1 if self._synced else 0
        v.max_ctas_per_sm = 1 if small_grid else 0  # overlapped with backward: one small CTA next to each persistent GEMM CTA
        return v

    def _fused_buffer(self, bucket: dict, key: str) -> DeviceArray:
        st = bucket["state"]
        if key not in st:
            st[key] = DeviceArray.zeros((bucket["shard"],), np.float32)  # the reference's moments start from 0
        return st[key]

    def _fused_rebind(self) -> None:
        """A parameter rebound since (load_state_dict) moves back into the arena."""
        f = self._fused
        base = f["P"].array.ptr
        for p, o in zip(self._parameters, self._offsets):
            if p.data.ptr != base + 4 * o:
                f["P"].tensor[o:o + p.size].copy_(p.data._buf.reshape(-1))
                p.data = DeviceArray(f["P"].tensor[o:o + p.size], p.shape, np.float32)

    def _fused_launch_bucket_async(self, bucket: dict) -> None:
        """Overlapped form: as soon as backward has enqueued the last gradient of a bucket, its exchange + update goes to a
        high-priority side stream (event wait -> cross-rank barrier -> fused kernel with a one-CTA-per-SM grid that co-resides
        with the persistent GEMM kernels of the layers still in backward).  Every rank issues the buckets in the same order."""
        import torch
        f = self._fused
        if f["side"] is None:
            f["side"] = torch.cuda.Stream(priority=-1)
        if f["scalars"] is None:
            f["scalars"] = self._step_scalars()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(f["side"]):
            f["side"].wait_event(ev)
            f["G"].barrier(1)
            # the bucket backward completes LAST (the first layers' parameters) has nothing left to hide behind: full grid
            last = bucket["order"] == len(f["buckets"]) - 1
            self._launch_fused(f["scalars"], 1.0 / f["world"], None, bucket, not last)
        bucket["launched"] = True

    def _fused_step(self, scalars, live) -> None:
        """barrier -> one kernel per bucket (in-switch gradient sum of this rank's slice, update, multicast of the new parameters)
        -> barrier; no separate collective and no host synchronisation.  Buckets already exchanged during backward
        (``overlap_grad_sync``) are only waited for."""
        import torch
        if any(p.grad is None for p in self._parameters):
            raise RuntimeError("fused data-parallel step: every parameter needs a gradient each step (parameters without one "
                               "are skipped by the reference, optimizers.py:157; set optimizer.fused_dp_step = False before "
                               "the first step for such models)")
        self._gather_grads_into_arena()
        self._fused_rebind()
        f = self._fused
        overlapped = any(b["launched"] for b in f["buckets"])
        if self.local_weight != 1.0 and not self._synced:
            if overlapped:
                raise RuntimeError("overlap_grad_sync needs equal shards (local_weight == 1.0)")
            self._arena *= float(self.local_weight)  # this rank's share of the global-batch mean (uneven shards)
        if overlapped:
            if self._synced:
                raise RuntimeError("overlap_grad_sync: gradients were synchronised by hand (clip_grad_norm / sync_grads) after "
                                   "buckets had already been exchanged during backward; use one or the other")
            for b in f["buckets"]:  # whatever backward did not complete goes behind the others on the side stream
                if not b["launched"]:
                    self._fused_launch_bucket_async(b)
            done = torch.cuda.Event()
            done.record(f["side"])
            torch.cuda.current_stream().wait_event(done)
        else:
            if scalars is None:
                scalars = self._step_scalars()
            f["G"].barrier(0)   # every rank's backward has written its gradient arena
            for b in f["buckets"]:
                self._launch_fused(scalars, 1.0 if self._synced else 1.0 / f["world"], live, b, False)
        f["P"].barrier(0)       # every replica of the parameters is complete before the next forward reads it
        self._fused_reset_buckets()

    def _fused_reset_buckets(self) -> None:
        if self._fused is not None:
            self._fused["scalars"] = None
            for b in self._fused["buckets"]:
                b["pending"], b["launched"] = len(b["members"]), False

    def _launch_fused(self, sc: list[float], scale: float, live, bucket: dict, small_grid: bool) -> None:
        raise NotImplementedError

    def fused_dp_note(self) -> str:
        if self._fused is None:
            return "off (" + (distributed.symmetric_memory_note() or "single rank / unsupported optimizer") + ")"
        nb = len(self._fused["buckets"])
        how = f", {nb} buckets overlapped with backward" if self.overlap_grad_sync and nb > 1 else ""
        if self._fused["use_multimem"]:
            return "multimem (NVLS in-switch reduction + multicast store)" + how
        return "peer loads / stores" + (" (2 ranks: cheaper than pulling both copies through the switch)" if self._fused["P"].multicast_ptr else " (no multicast support)") + how

    def _materialize_fused_state(self) -> None:
        """Checkpoints keep the reference layout ``{i: {"m": Tensor, "v": Tensor}}``: the slices are gathered from all ranks and
        cut at the parameter offsets (collective call)."""
        for b in self._fused["buckets"]:
            for key, shard in b["state"].items():
                full = distributed.all_gather(shard).reshape(-1)  # [world * shard] == the bucket's range [lo, hi)
                for i in b["members"]:
                    p, o = self._parameters[i], self._offsets[i] - b["lo"]
                    self._state[i][key] = Tensor(DeviceArray(full._buf[o:o + p.size].clone(), p.shape, np.float32))

    def _scatter_fused_state(self) -> None:
        for bk in self._fused["buckets"]:
            lo, hi = bk["off"], bk["off"] + bk["shard"]
            for key in self._fused_dp_keys or ():
                buf = self._fused_buffer(bk, key)
                buf.fill(0.0)
                for i in bk["members"]:
                    p, o = self._parameters[i], self._offsets[i]
                    st = self._state.get(i, {}).get(key)
                    a, b = max(o, lo), min(o + p.size, hi)
                    if st is None or a >= b:
                        continue
                    src = st.data if isinstance(st.data, DeviceArray) else DeviceArray.from_numpy(np.asarray(st.data, np.float32))
                    buf._buf[a - lo:b - lo].copy_(src._buf.reshape(-1)[a - o:b - o])

    # ---- overlapped data-parallel exchange -----------------------------------------------------
    def _dp_world(self) -> int:
        return distributed.world_size() if self._data_parallel else 1

    def _reset_buckets(self) -> None:
        if self._buckets is not None:
            self._bucket_pending = [len(m) for _, _, m in self._buckets]
            self._bucket_launched = [False] * len(self._buckets)
            self._bucket_work = []

    def _ensure_buckets(self) -> None:
        if self._buckets is None:
            self._buckets = plan_grad_buckets(self._offsets, [p.size for p in self._parameters], max(1, self.bucket_bytes // 4))
            self._bucket_of = {}
            for b, (_, _, members) in enumerate(self._buckets):
                for i in members:
                    self._bucket_of[i] = b
            self._reset_buckets()

    def _launch_bucket(self, b: int) -> None:
        lo, hi, _ = self._buckets[b]
        if not any(self._bucket_launched) and self.reserve_sms > 0:
            _lib.check(_lib.lib().cpt_tc_reserve_sms(int(self.reserve_sms)))
        self._bucket_launched[b] = True
        self._bucket_work.append(distributed.all_reduce_sum_async(self._arena._buf[lo:hi]))

    def _on_grad_ready(self, i: int, first: bool) -> None:
        """Called by ``Module.update_parameter_grad`` once parameter i's gradient has been written into its arena slot
        (``first``) or accumulated into it again (shared parameter)."""
        if not self.overlap_grad_sync or self._arena is None or self._dp_world() == 1 or graph.is_capturing():
            return
        if self._fused is not None:  # bucketed fused step on the side stream (the first step, which builds the arenas, is not overlapped)
            bk = self._fused["bucket_of"][i]
            if bk["launched"]:
                raise RuntimeError("overlap_grad_sync: a parameter received a second gradient after its bucket was exchanged "
                                   "(shared parameters need overlap_grad_sync = False)")
            if first:
                bk["pending"] -= 1
                if bk["pending"] == 0 and len(self._fused["buckets"]) > 1:
                    self._fused_launch_bucket_async(bk)
            return
        if self._fused_dp_keys is not None and self.fused_dp_step and not self._fused_checked:
            return  # the first step() decides between the fused and the classic exchange
        self._ensure_buckets()
        b = self._bucket_of[i]
        if not first and not self._bucket_launched[b]:
            return  # accumulated before the exchange: fine
        if self._bucket_launched[b]:
            raise RuntimeError("overlap_grad_sync: a parameter received a second gradient after its bucket was all-reduced "
                               "(shared parameters need overlap_grad_sync = False)")
        self._bucket_pending[b] -= 1
        if self._bucket_pending[b] == 0:
            self._launch_bucket(b)

    def _gather_grads_into_arena(self) -> None:
        """A gradient that was assigned by hand (not written into its slot by a layer) is copied into the arena so
        that the single all-reduce covers it."""
        for p in self._parameters:
            if p.grad is not None and p.grad_slot is not None and p.grad.data.ptr != p.grad_slot.ptr:
                p.grad_slot.copy_from(p.grad.data)
                p.grad = Tensor(p.grad_slot)

    def _state_array(self, i: int, key: str, like: Parameter) -> DeviceArray:
        st = self._state[i]
        if key not in st:  # the reference starts from python 0.0 (optimizers.py:164, 256, 261): a zero buffer is identical
            st[key] = Tensor(DeviceArray.zeros(like.shape, np.float32))
        elif not isinstance(st[key].data, DeviceArray):
            st[key] = st[key].to_device(like.device)
        return st[key].data

    def _table(self, keys: tuple[str, ...]) -> tuple[int, int, int]:
        """Builds/refreshes the device pointer table; returns (table ptr, entries, max elements)."""
        n = len(self._parameters)
        entries = (_lib.ParamEntry * max(n, 1))()
        max_n = 0
        for i, p in enumerate(self._parameters):
            if not isinstance(p.data, DeviceArray):
                raise TypeError("compyute_b200 optimizers update cuda parameters only (no CPU fallback).")
            e = entries[i]
            e.p = p.data.ptr
            if p.grad is None:
                e.n = 0  # skipped like `if p.grad is None: continue`
                continue
            e.g = p.grad.data.ptr
            e.m = self._state_array(i, keys[0], p).ptr if len(keys) > 0 else None
            e.v = self._state_array(i, keys[1], p).ptr if len(keys) > 1 else None
            e.n = p.size
            max_n = max(max_n, p.size)
        raw = bytes(entries)
        if raw != self._table_key:  # pointers are stable with the arena: the upload happens once, not per step
            host = np.frombuffer(raw, dtype=np.uint8).copy()
            self._table_dev = DeviceArray.from_numpy(host)
            self._table_key = raw
        return self._table_dev.ptr, n, max_n

    def _sync_grads(self) -> float:
        """Data-parallel exchange: one SUM all-reduce of the gradient arena; returns the 1/world scale."""
        world = distributed.world_size() if self._data_parallel else 1
        if world == 1:
            return 1.0
        if self._arena is None:
            raise RuntimeError("data-parallel step needs cuda parameters (gradient arena missing)")
        self._gather_grads_into_arena()
        if self.local_weight != 1.0:
            if self.overlap_grad_sync and self._buckets is not None and any(self._bucket_launched):
                raise RuntimeError("overlap_grad_sync needs equal shards (local_weight == 1.0)")
            self._arena *= float(self.local_weight)  # this rank's share of the global-batch mean (uneven shards)
        if self.overlap_grad_sync and self._buckets is not None and any(self._bucket_launched):
            for b in range(len(self._buckets)):  # whatever backward did not complete (e.g. parameters without gradient)
                if not self._bucket_launched[b]:
                    self._launch_bucket(b)
            for w in self._bucket_work:
                w.wait()  # the compute stream waits for NCCL's stream; no host sync
            if self.reserve_sms > 0:
                _lib.check(_lib.lib().cpt_tc_reserve_sms(0))
            self._reset_buckets()
        else:
            distributed.all_reduce_sum(self._arena)
        return 1.0 / world


class SGD(Optimizer):
    """optimizers.py:95-176"""

    def __init__(self, parameters: Optional[Iterable[Parameter]] = None, lr: float = 1e-3, momentum: float = 0.0,
                 nesterov: bool = False, weight_decay: float = 0.0) -> None:
        super().__init__(parameters, lr)
        self.momentum, self.nesterov, self.weight_decay = momentum, nesterov, weight_decay

    _fused_dp_keys = ("v",)

    def _step_scalars(self) -> list[float]:
        return [float(self.lr)]

    def _launch_fused(self, sc, scale, live, bucket, small_grid) -> None:
        vel = self._fused_buffer(bucket, "v").ptr if self.momentum > 0.0 else None
        view = self._fused_view(bucket, small_grid)
        _lib.check(_lib.lib().cpt_dp_sgd_step(ctypes.byref(view), vel, sc[0], float(self.momentum), int(self.nesterov),
                                              float(self.weight_decay), float(scale), live, stream_ptr()))

    def _launch(self, sc, live, scale) -> None:
        keys = ("v",) if self.momentum > 0.0 else ()
        table, n, max_n = self._table(keys)
        _lib.check(_lib.lib().cpt_sgd_step(table, n, max_n, sc[0], float(self.momentum), int(self.nesterov),
                                           float(self.weight_decay), float(scale), live, stream_ptr()))


class Adam(Optimizer):
    """optimizers.py:179-271"""

    _decoupled = 0
    _fused_dp_keys = ("m", "v")

    def __init__(self, parameters: Optional[Iterable[Parameter]] = None, lr: float = 1e-3, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0) -> None:
        super().__init__(parameters, lr)
        self.beta1, self.beta2, self.eps, self.weight_decay = beta1, beta2, eps, weight_decay

    def _launch_fused(self, sc, scale, live, bucket, small_grid) -> None:
        view = self._fused_view(bucket, small_grid)
        _lib.check(_lib.lib().cpt_dp_adam_step(ctypes.byref(view), self._fused_buffer(bucket, "m").ptr, self._fused_buffer(bucket, "v").ptr, sc[0],
                                               float(self.beta1), float(self.beta2), float(self.eps), float(self.weight_decay), sc[1],
                                               sc[2], float(scale), self._decoupled, live, stream_ptr()))

    def _step_scalars(self) -> list[float]:
        # python doubles, like optimizers.py:243-244
        return [float(self.lr), 1.0 - self.beta1 ** self.t, 1.0 - self.beta2 ** self.t]

    def _launch(self, sc, live, scale) -> None:
        table, n, max_n = self._table(("m", "v"))
        _lib.check(_lib.lib().cpt_adam_step(table, n, max_n, sc[0], float(self.beta1), float(self.beta2), float(self.eps),
                                            float(self.weight_decay), sc[1], sc[2], float(scale), self._decoupled, live,
                                            stream_ptr()))


class AdamW(Adam):
    """optimizers.py:274-362 (decoupled weight decay, default 1e-2)."""

    _decoupled = 1

    def __init__(self, parameters: Optional[Iterable[Parameter]] = None, lr: float = 1e-3, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 1e-2) -> None:
        super().__init__(parameters, lr, beta1, beta2, eps, weight_decay)


class NAdam(Optimizer):
    """optimizers.py:365-475 (Nesterov-accelerated Adam with the momentum-decay schedule)."""

    def __init__(self, parameters: Optional[Iterable[Parameter]] = None, lr: float = 2e-3, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0, momentum_decay: float = 4e-3) -> None:
        super().__init__(parameters, lr)
        self.beta1, self.beta2, self.eps, self.weight_decay = beta1, beta2, eps, weight_decay
        self.momentum_decay = momentum_decay
        self._mu_prod = 1.0

    def _step_scalars(self) -> list[float]:
        mu = self.beta1 * (1.0 - 0.5 * 0.96 ** (self.t * self.momentum_decay))  # python doubles, like :438-447
        mu_next = self.beta1 * (1.0 - 0.5 * 0.96 ** ((self.t + 1) * self.momentum_decay))
        self._mu_prod *= mu
        m_div = 1.0 - self._mu_prod * mu_next
        g_div = 1.0 - self._mu_prod
        v_div = 1.0 - self.beta2 ** self.t
        return [float(self.lr), m_div, v_div, mu, mu_next, g_div]  # order of the kernel's live-scalar layout

    def _launch(self, sc, live, scale) -> None:
        table, n, max_n = self._table(("m", "v"))
        _lib.check(_lib.lib().cpt_nadam_step(table, n, max_n, sc[0], float(self.beta1), float(self.beta2), float(self.eps),
                                             float(self.weight_decay), sc[3], sc[4], sc[1], sc[5], sc[2], float(scale), live,
                                             stream_ptr()))
