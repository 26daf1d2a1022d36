"""CUDA-graph capture of a train step (hard part 4: small-image configs are host-bound — ~80 launches of
a few microseconds each per step, against ~30 us of Python dispatch per launch).

    step = cp.graph.CapturedStep(train_step, optimizers=[opt])   # train_step(): fwd + loss + bwd + opt.step() on STATIC tensors
    for xb, tb in loader:
        x_static.data.upload(xb); t_static.data.upload(tb)         # new batch into the captured input buffers
        loss = step()                                              # one cudaGraphLaunch

What makes a captured step stay correct across replays:
  * parameters, optimizer moments and the gradient arena are updated in place (addresses are stable);
  * per-step optimizer scalars (lr, bias corrections) are read by the update kernel from a small device buffer that is
    refreshed before every replay, so LR schedulers and ``t`` keep working (Appendix A.17);
  * BatchNorm running statistics are updated in place while capturing (eager mode rebinds ``.data`` like the reference);
  * Dropout mixes a device-resident replay counter into its seed.
Host syncs (``.item()``, int64 label conversion) are not allowed inside the captured function.

Side effects of construction: ``fn`` runs eagerly ``warmup`` times before the capture (first-use initialisation, allocator
warm-up, optimizer state creation), i.e. ``warmup`` REAL optimizer updates on whatever the static input buffers hold — feed
real batches during the warm-up (what bench.py and the tests do) or pass ``restore_after_warmup=True``, which snapshots the
parameters, buffers and optimizer state (moments, ``t``, NAdam's running product) before the warm-up and puts them back
before the capture, so that the first replay is step 1 of the reference trace.

Lifetime: the captured kernels hold raw pointers into the process-wide scratch buffer (``tensors.workspace``) and into each
optimizer's device pointer table.  The CapturedStep keeps both alive (a later, larger ``workspace()`` request allocates a new
buffer for eager code and leaves the old one to the graph), and refuses to replay after ``Optimizer.load_state_dict`` /
``set_parameters`` replaced the state tensors the graph updates — re-capture in that case.
"""

from __future__ import annotations

from typing import Callable, Iterable

import numpy as np

_capturing = False
_replay_counter = None  # DeviceArray uint64[1], bumped before every replay (Dropout's live seed)


def is_capturing() -> bool:
    return _capturing


def replay_counter():
    """Device uint64 counter mixed into Dropout seeds inside captured steps."""
    global _replay_counter
    if _replay_counter is None:
        from .tensors import DeviceArray
        _replay_counter = DeviceArray.zeros((1,), np.uint64)
    return _replay_counter


class CapturedStep:
    def __init__(self, fn: Callable[[], object], optimizers: Iterable = (), warmup: int = 3, restore_after_warmup: bool = False,
                 modules: Iterable = ()) -> None:
        import torch
        global _capturing
        from . import tensors
        self._torch = torch
        self.optimizers = list(optimizers)
        self._count = 0
        replay_counter()
        for o in self.optimizers:
            o._live_buffer()  # created before the capture: an allocation + memset inside it would be replayed every time
        snap = self._snapshot(modules) if restore_after_warmup else None
        for i in range(max(1, warmup)):  # eager: first-use initialisation, allocator warm-up, optimizer state creation
            fn()
            if snap is not None and i == 0:
                snap = self._snapshot_late(snap)  # optimizer moments exist only after the first step
        torch.cuda.synchronize()
        if snap is not None:
            self._restore(snap)
        self.graph = torch.cuda.CUDAGraph()
        _capturing = True
        try:
            with torch.cuda.graph(self.graph):
                self.out = fn()
        finally:
            _capturing = False
        # what the captured kernels point into: kept alive for the life of the graph (a later, larger workspace()
        # request used to hand the old scratch back to the caching allocator while replays still wrote into it)
        self._pinned = (tensors.current_workspace(), [o._table_dev for o in self.optimizers], [o._arena for o in self.optimizers])
        self._table_keys = [o._table_key for o in self.optimizers]
        # BatchNorm running statistics are updated in place by the graph: an eager training forward rebinds ``.data``
        # (normalizations.py:92-97), after which the graph would keep updating the orphaned arrays
        self._buffers = [(b, b.data.ptr) for m in modules for b in m.get_buffers() if hasattr(b.data, "ptr")]

    # ---- warm-up without side effects (optional)
    def _snapshot(self, modules):
        params, seen = [], set()
        for o in self.optimizers:
            for p in o._parameters:
                if id(p) not in seen:
                    seen.add(id(p)); params.append(p)
        bufs = [b for m in modules for b in m.get_buffers()]
        pre = {(oi, i, k): st.data.copy() for oi, o in enumerate(self.optimizers) for i in o._state for k, st in o._state[i].items()
               if hasattr(st.data, "copy_from")}
        return {"params": [(p, p.data.copy()) for p in params], "bufs": [(b, b.data.copy()) for b in bufs],
                "opt": [dict(t=o.t, mu=getattr(o, "_mu_prod", None)) for o in self.optimizers], "pre": pre, "state": None}

    def _snapshot_late(self, snap):
        # moments that did not exist before are created (zero-filled) by the first step: they are zeroed again in place
        snap["state"] = [[(i, k, st) for i in sorted(o._state) for k, st in o._state[i].items()] for o in self.optimizers]
        return snap

    def _restore(self, snap) -> None:
        for t, saved in snap["params"] + snap["bufs"]:
            t.data.copy_from(saved)
        for oi, (o, meta, state) in enumerate(zip(self.optimizers, snap["opt"], snap["state"] or [[] for _ in self.optimizers])):
            o.t = meta["t"]
            if meta["mu"] is not None:
                o._mu_prod = meta["mu"]
            for i, k, st in state:
                if (oi, i, k) in snap["pre"]:
                    st.data.copy_from(snap["pre"][(oi, i, k)])
                else:
                    st.data.fill(0.0)  # the reference's moments start from 0 (optimizers.py:164, 256, 261)

    def __call__(self):
        from . import _lib
        from .tensors import stream_ptr
        for o, key in zip(self.optimizers, self._table_keys):
            if o._table_key is not key:
                raise RuntimeError("CapturedStep: the optimizer's parameters / state tensors were replaced after the capture "
                                   "(load_state_dict, set_parameters): the graph would update the old buffers — re-capture")
        for b, ptr in self._buffers:
            if b.data.ptr != ptr:
                raise RuntimeError("CapturedStep: a module buffer was rebound after the capture (an eager training forward between "
                                   "replays?): the graph updates the old array — re-capture")
        self._count += 1
        _lib.check(_lib.lib().cpt_set_u64(replay_counter().ptr, self._count, stream_ptr()))
        for o in self.optimizers:
            o.upload_live_scalars()  # scalars of step t -> device, then t += 1 (what step() does in eager mode)
        self.graph.replay()
        return self.out
