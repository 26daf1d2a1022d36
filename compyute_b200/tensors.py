"""Tensor and device array.  Mirrors compyute/tensors.py:76-689 for the surface the CNN hot path touches.

``Tensor.data`` is a NumPy array on ``cpu`` and a ``DeviceArray`` on ``cuda``.  A ``DeviceArray`` is only a typed,
shaped handle on device memory (owned by torch's caching allocator — plumbing); all arithmetic on it goes
through libcompyute_b200's kernels via the C ABI.
"""

from __future__ import annotations

import ctypes
from typing import Any, Optional

import numpy as np

from . import _lib
from .backend import Device, DeviceError, cpu, cuda, select_device

__all__ = ["Tensor", "DeviceArray", "ShapeError", "tensor", "stream_ptr", "workspace"]

ShapeLike = tuple[int, ...]


class ShapeError(Exception):
    """Incompatible tensor shapes (tensors.py:38)."""


_torch = None


def _t():
    global _torch
    if _torch is None:
        import torch
        _torch = torch
    return _torch


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream: every kernel of the path is ordered on it."""
    return _t().cuda.current_stream().cuda_stream


_NP2TORCH = {"float32": "float32", "int32": "int32", "int64": "int64", "int8": "int8", "uint8": "uint8", "bool": "bool",
             "float64": "float64", "bfloat16": "bfloat16", "uint64": "uint64"}


class DeviceArray:
    """C-contiguous array in HBM.  ``ptr`` is what the C ABI receives (== CuPy's ``arr.data.ptr``)."""

    __slots__ = ("_buf", "shape", "dtype", "cl", "stats")

    def __init__(self, buf, shape: ShapeLike, dtype: np.dtype):
        # cl: optional (mode, DeviceArray, chan_sum | None) — the channels-last low-precision copy of this array that its
        # PRODUCER kernel wrote alongside it (cpt_bn_act_*_cl), consumed by the next tensor-core convolution instead of a
        # staging pass; every in-place mutation drops it
        self.cl = self.stats = None
        # stats: optional (slots array, n_slots, conv bias | None) — per-channel Σ / Σ² partials of this array left by the
        # convolution epilogue that produced it, consumed by a BatchNorm forward instead of its statistics pass
        self.stats = None
        self._buf = buf  # torch tensor owning the memory (1-D or any shape, contiguous)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)

    # -- creation
    @staticmethod
    def empty(shape: ShapeLike, dtype=np.float32) -> "DeviceArray":
        torch = _t()
        if not torch.cuda.is_available():
            raise DeviceError("CUDA device not available (compyute_b200 has no CPU fallback).")
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        buf = torch.empty(shape, dtype=getattr(torch, _NP2TORCH[np.dtype(dtype).name]), device=f"cuda:{cuda.index}")
        return DeviceArray(buf, shape, dtype)

    @staticmethod
    def zeros(shape: ShapeLike, dtype=np.float32) -> "DeviceArray":
        a = DeviceArray.empty(shape, dtype)
        a._buf.zero_()  # cudaMemsetAsync
        return a

    @staticmethod
    def from_numpy(a: np.ndarray) -> "DeviceArray":
        torch = _t()
        a = np.ascontiguousarray(a)
        out = DeviceArray.empty(a.shape, a.dtype)
        src = torch.from_numpy(a.reshape(-1).copy() if a.ndim == 0 else a)
        out._buf.copy_(src.reshape(out._buf.shape), non_blocking=False)  # cudaMemcpy H2D
        return out

    # -- attributes
    @property
    def ptr(self) -> int:
        return self._buf.data_ptr()

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def strides(self) -> tuple[int, ...]:
        s, acc = [], self.dtype.itemsize
        for d in reversed(self.shape):
            s.append(acc)
            acc *= d
        return tuple(reversed(s))

    @property
    def __cuda_array_interface__(self) -> dict:
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    # -- views / copies
    def reshape(self, *shape) -> "DeviceArray":
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = -int(np.prod(shape, dtype=np.int64))
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ShapeError(f"cannot reshape {self.shape} into {shape}")
        return DeviceArray(self._buf, shape, self.dtype)  # zero-copy view (C-contiguous)

    def copy(self) -> "DeviceArray":
        return DeviceArray(self._buf.clone(), self.shape, self.dtype)  # cudaMemcpyAsync D2D

    def copy_from(self, other: "DeviceArray") -> None:
        self.cl = self.stats = None
        if other.size != self.size or other.dtype != self.dtype:
            raise ShapeError(f"copy_from: {other.shape}/{other.dtype} into {self.shape}/{self.dtype}")
        self._buf.view(-1).copy_(other._buf.view(-1))

    def upload(self, a) -> None:
        self.cl = self.stats = None
        """In-place H2D copy into this buffer (keeps the address: used to feed CUDA-graph-captured steps)."""
        torch = _t()
        if isinstance(a, DeviceArray):
            return self.copy_from(a)
        if hasattr(a, "to_numpy"):
            a = a.data if isinstance(a.data, DeviceArray) else a.to_numpy()
            if isinstance(a, DeviceArray):
                return self.copy_from(a)
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.size != self.size:
            raise ShapeError(f"upload: {a.shape} into {self.shape}")
        self._buf.view(-1).copy_(torch.from_numpy(a).view(-1), non_blocking=True)

    def numpy(self) -> np.ndarray:
        return self._buf.detach().cpu().numpy().reshape(self.shape)  # cudaMemcpy D2H (synchronising)

    def item(self):
        return self.numpy().reshape(-1)[0].item()

    # -- arithmetic through the C ABI (fp32 only)
    def _f32(self, who: str) -> None:
        if self.dtype != np.float32:
            raise TypeError(f"{who}: only float32 device arrays are supported, got {self.dtype}")

    def __iadd__(self, other) -> "DeviceArray":
        self._f32("+=")
        self.cl = self.stats = None
        if isinstance(other, DeviceArray):
            if other.size != self.size:
                raise ShapeError(f"+=: shapes {self.shape} and {other.shape} differ (no broadcasting on device)")
            _lib.check(_lib.lib().cpt_add_inplace(self.ptr, other.ptr, self.size, stream_ptr()))
            return self
        raise TypeError(f"+=: unsupported operand {type(other)}")

    def fill(self, value: float) -> None:
        self.cl = self.stats = None
        self._f32("fill")
        _lib.check(_lib.lib().cpt_fill(self.ptr, float(value), self.size, stream_ptr()))

    def scaled(self, alpha: float) -> "DeviceArray":
        self._f32("scale")
        out = DeviceArray.empty(self.shape, self.dtype)
        _lib.check(_lib.lib().cpt_axpby(out.ptr, self.ptr, float(alpha), 0, self.size, stream_ptr()))
        return out

    def sum(self) -> float:
        self._f32("sum")
        out = DeviceArray.empty((1,), np.float32)
        _lib.check(_lib.lib().cpt_sum(self.ptr, self.size, out.ptr, stream_ptr()))
        return out.item()

    def has_nan(self) -> bool:
        """``is_nan(x).any().item()`` (module.py:332) — one reduction kernel + one 4-byte D2H."""
        self._f32("has_nan")
        flag = DeviceArray.zeros((1,), np.int32)
        _lib.check(_lib.lib().cpt_isnan_flag(self.ptr, self.size, flag.ptr, stream_ptr()))
        return bool(flag.item())

    def __repr__(self) -> str:
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, ptr=0x{self.ptr:x})"

    def __reduce__(self):
        """Pickling goes through a host copy (``cp.save(model.get_state_dict())`` keeps working, utils.py:44-73);
        unpickling uploads to this process's CUDA device, like unpickling a CuPy array does."""
        return (DeviceArray.from_numpy, (self.numpy(),))


# -- per-process scratch: one growable buffer, safe because every op of the path is ordered on one stream
_ws: Optional[DeviceArray] = None


def workspace(nbytes: int) -> tuple[int, int]:
    """Returns (ptr, nbytes) of a scratch buffer of at least ``nbytes``."""
    global _ws
    nbytes = max(int(nbytes), 256)
    if _ws is None or _ws.nbytes < nbytes:
        _ws = DeviceArray.empty(((nbytes + (1 << 20) - 1) // (1 << 20) * (1 << 20),), np.uint8)
    return _ws.ptr, _ws.nbytes


def _device_of(data: Any) -> Device:
    return cuda if isinstance(data, DeviceArray) else cpu


class Tensor:
    """Multi-dimensional array with a gradient slot (tensors.py:76-95)."""

    def __init__(self, data: Any) -> None:
        if isinstance(data, Tensor):
            data = data.data
        if not isinstance(data, (np.ndarray, DeviceArray)):
            data = np.asarray(data)
        self.data = data
        self.grad: Optional[Tensor] = None
        self._iterator = 0

    # ---- properties (tensors.py:100-150)
    @property
    def device(self) -> Device:
        return _device_of(self.data)

    @property
    def dtype(self) -> np.dtype:
        return self.data.dtype

    @property
    def shape(self) -> ShapeLike:
        return tuple(self.data.shape)

    @property
    def ndim(self) -> int:
        return self.data.ndim

    @property
    def size(self) -> int:
        return int(self.data.size)

    @property
    def strides(self) -> tuple[int, ...]:
        return tuple(self.data.strides)

    @property
    def ptr(self) -> int:
        """Identity of the underlying array (tensors.py:138-140) — optimizers de-duplicate parameters by it."""
        return id(self.data)

    @property
    def T(self) -> "Tensor":
        if self.device is cuda:
            raise NotImplementedError("transpose of device tensors is not part of the CNN hot path")
        return Tensor(np.swapaxes(self.data, -1, -2))

    def __bool__(self) -> bool:  # tensors.py:305-306: `if b:` means `b is not None`
        return True

    def __len__(self) -> int:
        return self.shape[0]

    def __repr__(self) -> str:
        if self.device is cuda:
            return f"Tensor({self.to_numpy()!r}, device=cuda)"
        return f"Tensor({self.data!r})"

    # ---- device / dtype movement (tensors.py:312-401)
    def to_device(self, device: Device) -> "Tensor":
        if device == self.device or (device.t == self.device.t):
            return self
        if device.t == "cuda":
            arr = self.data
            if arr.dtype == np.float64:
                arr = arr.astype(np.float32)
            return Tensor(DeviceArray.from_numpy(arr))
        return Tensor(self.data.numpy())

    def ito_device(self, device: Device) -> None:
        if device.t == self.device.t:
            return
        self.data = self.to_device(device).data
        if self.grad is not None:
            self.grad = self.grad.to_device(device)

    def to_numpy(self) -> np.ndarray:
        return self.data.numpy() if isinstance(self.data, DeviceArray) else self.data

    def item(self):
        return self.data.item()

    def copy(self) -> "Tensor":
        return Tensor(self.data.copy())

    def view(self, shape: ShapeLike) -> "Tensor":
        return Tensor(self.data.reshape(shape))

    def to_contiguous(self) -> "Tensor":
        return self if isinstance(self.data, DeviceArray) else Tensor(np.ascontiguousarray(self.data))

    # ---- arithmetic used by the callers of the hot path
    def __iadd__(self, other: "Tensor") -> "Tensor":
        o = other.data if isinstance(other, Tensor) else other
        if isinstance(self.data, DeviceArray):
            self.data += o
        else:
            self.data += o
        return self

    def __add__(self, other: "Tensor") -> "Tensor":
        out = self.copy()
        out += other
        return out

    def sum(self) -> float:
        return self.data.sum() if isinstance(self.data, DeviceArray) else float(self.data.sum())


def tensor(data: Any, device: Optional[Device] = None, dtype=None) -> Tensor:
    """``compyute.tensor`` (tensors.py:44-73): host data → Tensor on ``device`` (default: context / cpu)."""
    arr = np.asarray(data, dtype=dtype)
    if arr.dtype == np.float64 and dtype is None:
        arr = arr.astype(np.float32)
    t = Tensor(arr)
    return t.to_device(select_device(device))


def require_cuda(*tensors: Optional[Tensor]) -> None:
    """Hot-path Functions accept device tensors only."""
    for t in tensors:
        if t is not None and not isinstance(t.data, DeviceArray):
            raise DeviceError("compyute_b200 functions run on cuda tensors only (no CPU fallback); "
                              "move the tensor with .to_device(cuda).")


def f32ptr(t: Optional[Tensor]) -> Optional[int]:
    if t is None:
        return None
    if t.data.dtype != np.float32:
        raise TypeError(f"expected float32 tensor, got {t.data.dtype}")
    return t.data.ptr
