"""``compyute.random`` for ``cuda`` tensors (compyute/random/random.py:25-204).

The device generator is counter-based (csrc/tensor_ops.cu ``random_kernel``): reproducible under ``set_seed`` and
statistically equivalent to, not stream-identical with, NumPy's generator.  ``cpu`` tensors use NumPy like the reference.
"""

from __future__ import annotations

from contextlib import contextmanager
from typing import Optional

import numpy as np

from . import device_ops as D
from .backend import Device, select_device
from .tensors import DeviceArray, Tensor

__all__ = ["set_seed", "seed", "random", "normal", "uniform", "uniform_int", "permutation", "bernoulli", "shuffle"]


def _shape(shape):
    return tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))


def set_seed(value: Optional[int] = None) -> None:
    """random.py:25-36: seeds the host (NumPy) and the device generator."""
    np.random.seed(value)
    D.set_seed(value)


@contextmanager
def seed(value: int):
    """random.py:39-51: ``with seed(42): ...`` / ``@seed(42)`` — seeded inside, reset afterwards."""
    set_seed(value)
    try:
        yield
    finally:
        set_seed()


def _draw(shape, kind, p0, p1, dtype, device, host):
    if select_device(device).t == "cuda":
        a = D.random_fill(_shape(shape), kind, p0, p1)
        return Tensor(a if np.dtype(dtype) == np.float32 else D.astype(a, dtype))
    return Tensor(host().astype(dtype))


def random(shape, *, device: Optional[Device] = None, dtype=np.float32) -> Tensor:
    """random.py:54-80: uniform [0, 1)"""
    return _draw(shape, 0, 0.0, 1.0, dtype, device, lambda: np.random.random(_shape(shape)))


def normal(shape, mean: float = 0.0, std: float = 1.0, *, device: Optional[Device] = None, dtype=np.float32) -> Tensor:
    """random.py:83-115"""
    return _draw(shape, 1, mean, std, dtype, device, lambda: np.random.normal(mean, std, _shape(shape)))


def uniform(shape, low: float = 0.0, high: float = 1.0, *, device: Optional[Device] = None, dtype=np.float32) -> Tensor:
    """random.py:118-150"""
    return _draw(shape, 0, low, high, dtype, device, lambda: np.random.uniform(low, high, _shape(shape)))


def uniform_int(shape, low: int, high: int, *, device: Optional[Device] = None, dtype=np.int32) -> Tensor:
    """random.py:153-184: integers in [low, high) (exact below 2^24)"""
    return _draw(shape, 2, low, high, dtype, device, lambda: np.random.randint(low, high, _shape(shape)))


def permutation(n: int, *, device: Optional[Device] = None) -> Tensor:
    """random.py:187-205: drawn on the host (n int32 indices), uploaded when ``device`` is cuda."""
    p = np.random.permutation(int(n)).astype(np.int32)
    return Tensor(DeviceArray.from_numpy(p)) if select_device(device).t == "cuda" else Tensor(p)


def bernoulli(p: float, shape, *, device: Optional[Device] = None) -> Tensor:
    """random.py:236-265: random() < p"""
    return random(shape, device=device) < p


def shuffle(x: Tensor) -> tuple[Tensor, Tensor]:
    """random.py:268-285: rows of x in a random order + the permutation used"""
    idx = permutation(x.shape[0], device=x.device)
    return x[idx], idx
