"""Input pipeline.

``Dataloader`` mirrors compyute/nn/utils/dataloaders.py:18-69 (same constructor, ``__call__`` yields tuples of batch
Tensors on ``device``, ``__len__``).  It is the H2D boundary of a train step and the natural data-parallel shard point
(SURVEY §8e, f3).  Two paths:

* data tensors already on ``cuda`` (the dataset fits in HBM — 180 GB): the shuffled batch is GATHERED ON THE DEVICE
  (``cpt_gather_rows`` with the batch's index vector, a few KB of H2D per step); nothing else crosses PCIe.
* data on the host: batches are gathered on the host into PINNED staging buffers and uploaded with an asynchronous copy on
  a side stream one batch ahead of the consumer.  A pinned slot is rewritten only after the H2D copy that last read it has
  finished (per-slot event, waited on the HOST: a device-side ``wait_event`` does not stop the host from running ahead
  when the consumer's host time per step is tiny, e.g. CUDA-graph replay).

With ``shard=True`` every rank takes its contiguous slice of each global batch (``distributed.shard_bounds``).
"""

from __future__ import annotations

from typing import Iterator, Optional

import numpy as np

from ... import distributed
from ...backend import Device, cpu
from ...tensors import DeviceArray, Tensor

__all__ = ["Dataloader"]


class Dataloader:
    def __init__(self, data: tuple[Tensor, ...], batch_size: int = 1, device: Device = cpu, shuffle_data: bool = True,
                 drop_remaining: bool = False, shard: bool = False) -> None:
        self.data = data
        self._n = len(self.data[0])
        self.batch_size = min(batch_size, self._n)
        self.device = device
        self.shuffle = shuffle_data
        self.shard = shard
        self._additional_batch = not drop_remaining and self._n % self.batch_size > 0
        self._pinned = None  # two sets of pinned staging buffers (double buffering)
        self._slot_events = [None, None]  # H2D completion of the copy that last read each pinned slot

    def __len__(self) -> int:
        return max(1, self._n // self.batch_size + self._additional_batch)

    def _host_batch(self, idx: np.ndarray) -> list[np.ndarray]:
        if self.shard:
            lo, hi = distributed.shard_bounds(len(idx))
            idx = idx[lo:hi]
        out = []
        for t in self.data:
            a = t.to_numpy()[idx]
            if a.dtype == np.int64:  # labels are int32 on the device
                a = a.astype(np.int32)
            elif a.dtype == np.float64:
                a = a.astype(np.float32)
            out.append(np.ascontiguousarray(a))
        return out

    def _device_batches(self, batches) -> Iterator[tuple[Tensor, ...]]:
        """Device-resident dataset: x[idx] on the device (dataloaders.py:65-66 does the same fancy index on CuPy arrays)."""
        from ... import device_ops as D
        for b in batches:
            if self.shard:
                lo, hi = distributed.shard_bounds(len(b))
                b = b[lo:hi]
            idx = DeviceArray.from_numpy(np.ascontiguousarray(b, dtype=np.int32))
            out = []
            for t in self.data:
                a = D.getitem(t.data, idx)
                if a.dtype == np.int64:  # labels are int32 on the device
                    a = D.astype(a, np.int32)
                out.append(Tensor(a))
            yield tuple(out)

    def __call__(self) -> Iterator[tuple[Tensor, ...]]:
        # same index stream as the reference: numpy's legacy global RNG (random.py permutation)
        idx = np.random.permutation(self._n) if self.shuffle else np.arange(self._n, dtype=np.int64)
        batches = [idx[i * self.batch_size:(i + 1) * self.batch_size] for i in range(len(self))]
        if self.device.t != "cuda":
            for b in batches:
                yield tuple(Tensor(a) for a in self._host_batch(b))
            return
        import torch
        if all(isinstance(t.data, DeviceArray) for t in self.data):
            yield from self._device_batches(batches)
            return
        side = torch.cuda.Stream()
        main = torch.cuda.current_stream()

        def upload(b, slot):
            host = self._host_batch(b)
            if self._pinned is None:
                self._pinned = [None, None]
            pins = self._pinned[slot]
            if pins is None or any(p.numel() < h.size or p.dtype != torch.from_numpy(h).dtype for p, h in zip(pins, host)):
                cap = [max(h.size, (self.batch_size * int(np.prod(h.shape[1:], dtype=np.int64)))) for h in host]
                pins = [torch.empty(c, dtype=torch.from_numpy(h).dtype).pin_memory() for c, h in zip(cap, host)]
                self._pinned[slot] = pins
            if self._slot_events[slot] is not None:
                self._slot_events[slot].synchronize()  # host-side: the previous H2D out of this slot's pinned buffers is done
            devs = []
            with torch.cuda.stream(side):
                for p, h in zip(pins, host):
                    p[:h.size].copy_(torch.from_numpy(h).view(-1))           # host gather -> pinned
                    d = torch.empty(h.shape, dtype=p.dtype, device="cuda")
                    d.copy_(p[:h.size].view(h.shape), non_blocking=True)       # async H2D on the side stream
                    devs.append((d, h))
                ev = torch.cuda.Event()
                ev.record(side)
            self._slot_events[slot] = ev
            return devs, ev

        pending = upload(batches[0], 0) if batches else None
        for i in range(len(batches)):
            devs, ev = pending
            if i + 1 < len(batches):
                pending = upload(batches[i + 1], (i + 1) & 1)  # next batch is in flight while this one is consumed
            main.wait_event(ev)
            out = []
            for d, h in devs:
                d.record_stream(main)
                out.append(Tensor(DeviceArray(d, h.shape, h.dtype)))
            yield tuple(out)
