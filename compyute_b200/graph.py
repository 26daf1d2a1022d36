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
    def __init__(self, fn: Callable[[], object], optimizers: Iterable = (), warmup: int = 3) -> None:
        import torch
        global _capturing
        self._torch = torch
        self.optimizers = list(optimizers)
        self._count = 0
        replay_counter()
        for o in self.optimizers:
            o._live_buffer()  # created before the capture: an allocation + memset inside it would be replayed every time
        for _ in range(max(1, warmup)):  # eager: first-use initialisation, allocator warm-up, optimizer state creation
            fn()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        _capturing = True
        try:
            with torch.cuda.graph(self.graph):
                self.out = fn()
        finally:
            _capturing = False

    def __call__(self):
        from . import _lib
        from .tensors import stream_ptr
        self._count += 1
        _lib.check(_lib.lib().cpt_set_u64(replay_counter().ptr, self._count, stream_ptr()))
        for o in self.optimizers:
            o.upload_live_scalars()  # scalars of step t -> device, then t += 1 (what step() does in eager mode)
        self.graph.replay()
        return self.out
