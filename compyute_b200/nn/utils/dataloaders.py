"""Input pipeline.

``Dataloader`` mirrors compyute/nn/utils/dataloaders.py:18-69 (same constructor, ``__call__`` yields tuples of batch
Tensors on ``device``, ``__len__``).  It is the H2D boundary of a train step and the natural data-parallel shard point
(SURVEY §8e, f3): batches are gathered on the host into PINNED staging buffers and uploaded with an asynchronous copy on
a side stream one batch ahead of the consumer; with ``shard=True`` every rank takes its contiguous slice of each global
batch (``distributed.shard_bounds``).
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

    def __call__(self) -> Iterator[tuple[Tensor, ...]]:
        # same index stream as the reference: numpy's legacy global RNG (random.py permutation)
        idx = np.random.permutation(self._n) if self.shuffle else np.arange(self._n, dtype=np.int64)
        batches = [idx[i * self.batch_size:(i + 1) * self.batch_size] for i in range(len(self))]
        if self.device.t != "cuda":
            for b in batches:
                yield tuple(Tensor(a) for a in self._host_batch(b))
            return
        import torch
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
            devs = []
            with torch.cuda.stream(side):
                for p, h in zip(pins, host):
                    p[:h.size].copy_(torch.from_numpy(h).view(-1))           # host gather -> pinned
                    d = torch.empty(h.shape, dtype=p.dtype, device="cuda")
                    d.copy_(p[:h.size].view(h.shape), non_blocking=True)       # async H2D on the side stream
                    devs.append((d, h))
                ev = torch.cuda.Event()
                ev.record(side)
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
