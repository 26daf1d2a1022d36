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

    @staticmethod
    def from_cuda_array_interface(obj) -> "DeviceArray":
        """Adopts foreign device memory WITHOUT a copy: any object exporting ``__cuda_array_interface__`` (a CuPy ndarray — what
        the reference's ``cuda`` tensors hold, backend.py:169-173 —, a Numba device array, a torch tensor).  The array must be
        C-contiguous and live on this process's device; the exporting object is kept alive by the returned handle."""
        torch = _t()
        if isinstance(obj, DeviceArray):
            return obj
        cai = obj.__cuda_array_interface__
        shape, dtype = tuple(int(s) for s in cai["shape"]), np.dtype(cai["typestr"])
        strides = cai.get("strides")
        if strides is not None:
            expect, acc = [], dtype.itemsize
            for d in reversed(shape):
                expect.append(acc); acc *= d
            if tuple(strides) != tuple(reversed(expect)):
                raise ShapeError("only C-contiguous device arrays can be adopted (call ascontiguousarray on the producer side)")
        if dtype.name not in _NP2TORCH:
            raise TypeError(f"unsupported device array dtype {dtype}")
        buf = obj if isinstance(obj, torch.Tensor) else torch.as_tensor(obj, device=f"cuda:{cuda.index}")  # zero-copy
        if buf.data_ptr() != int(cai["data"][0]):
            raise DeviceError("device array could not be adopted in place (different device?)")
        return DeviceArray(buf, shape, dtype)

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

    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        """``numpy.asarray(device_array)``: an explicit host copy (D2H, synchronising) — without it NumPy would fall back to
        the sequence protocol and issue one indexing launch per element."""
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    # -- arithmetic through the C ABI (fp32 only)
    def _f32(self, who: str) -> None:
        if self.dtype != np.float32:
            raise TypeError(f"{who}: only float32 device arrays are supported, got {self.dtype}")

    def fill(self, value: float) -> None:
        self.cl = self.stats = None
        self._f32("fill")
        _lib.check(_lib.lib().cpt_fill(self.ptr, float(value), self.size, stream_ptr()))

    def scaled(self, alpha: float) -> "DeviceArray":
        self._f32("scale")
        out = DeviceArray.empty(self.shape, self.dtype)
        _lib.check(_lib.lib().cpt_axpby(out.ptr, self.ptr, float(alpha), 0, self.size, stream_ptr()))
        return out

    # -- NumPy-style operator surface: what `Tensor.data <op> other` reaches on a CuPy array in the reference
    # (tensors.py:196-292); kernels in csrc/tensor_ops.cu, host-side shape logic in device_ops.py
    def _bin(self, op, other, reverse=False):
        from . import device_ops as D
        if other is None and op == "add" and reverse:  # tensors.py:199-201: None + grad
            return self.copy()
        return D.binary(op, self, other, reverse=reverse)

    def _ibin(self, op, other):
        from . import device_ops as D
        self.cl = self.stats = None
        if op == "add" and isinstance(other, DeviceArray) and other.shape == self.shape and other.dtype == self.dtype == np.float32:
            _lib.check(_lib.lib().cpt_add_inplace(self.ptr, other.ptr, self.size, stream_ptr()))  # residual / grad accumulation
            return self
        return D.binary(op, self, other, out=self)

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __iadd__(self, o): return self._ibin("add", o)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __isub__(self, o): return self._ibin("sub", o)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __imul__(self, o): return self._ibin("mul", o)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __itruediv__(self, o): return self._ibin("div", o)
    def __floordiv__(self, o): return self._bin("floordiv", o)
    def __rfloordiv__(self, o): return self._bin("floordiv", o, True)
    def __ifloordiv__(self, o): return self._ibin("floordiv", o)
    def __pow__(self, o): return self._bin("pow", o)
    def __rpow__(self, o): return self._bin("pow", o, True)
    def __ipow__(self, o): return self._ibin("pow", o)
    def __mod__(self, o): return self._bin("mod", o)
    def __rmod__(self, o): return self._bin("mod", o, True)
    def __imod__(self, o): return self._ibin("mod", o)
    def __lt__(self, o): return self._bin("lt", o)
    def __gt__(self, o): return self._bin("gt", o)
    def __le__(self, o): return self._bin("le", o)
    def __ge__(self, o): return self._bin("ge", o)
    def __eq__(self, o): return self._bin("eq", o)  # noqa: elementwise like NumPy
    def __ne__(self, o): return self._bin("ne", o)
    __hash__ = object.__hash__

    def __neg__(self):
        from . import device_ops as D
        return D.unary("neg", self)

    def __abs__(self):
        from . import device_ops as D
        return D.unary("abs", self)

    def __invert__(self):
        from . import device_ops as D
        if self.dtype != np.bool_:
            raise TypeError("~ is defined for bool device arrays only")
        return D.logic("not", self)

    def __and__(self, o):
        from . import device_ops as D
        return D.logic("and", self, o)

    def __or__(self, o):
        from . import device_ops as D
        return D.logic("or", self, o)

    def __xor__(self, o):
        from . import device_ops as D
        return D.logic("xor", self, o)

    def __matmul__(self, o):
        from . import device_ops as D
        return D.matmul(self, o)

    def __getitem__(self, key):
        from . import device_ops as D
        return D.getitem(self, key)

    def __setitem__(self, key, value) -> None:
        from . import device_ops as D
        D.setitem(self, key, value)

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def astype(self, dtype, copy: bool = True) -> "DeviceArray":
        from . import device_ops as D
        return self if (not copy and np.dtype(dtype) == self.dtype) else D.astype(self, dtype)

    def tolist(self) -> list:
        return self.numpy().tolist()

    def _red(self, op, dim, keepdims):
        from . import device_ops as D
        return D.reduce(op, self, dim, keepdims)

    def sum(self, dim=None, keepdims: bool = False) -> "DeviceArray":
        from . import device_ops as D
        return D.sum_(self, dim, keepdims)

    def mean(self, dim=None, keepdims: bool = False) -> "DeviceArray":
        from . import device_ops as D
        return D.mean(self, dim, keepdims)

    def var(self, dim=None, ddof: int = 0, keepdims: bool = False) -> "DeviceArray":
        from . import device_ops as D
        return D.var(self, dim, ddof, keepdims)

    def std(self, dim=None, keepdims: bool = False) -> "DeviceArray":
        from . import device_ops as D
        return D.std(self, dim, keepdims)

    def max(self, dim=None, keepdims: bool = False): return self._red("max", dim, keepdims)
    def min(self, dim=None, keepdims: bool = False): return self._red("min", dim, keepdims)
    def prod(self, dim=None, keepdims: bool = False): return self._red("prod", dim, keepdims)
    def any(self, dim=None, keepdims: bool = False): return self._red("any", dim, keepdims)
    def all(self, dim=None, keepdims: bool = False): return self._red("all", dim, keepdims)

    def argmax(self, dim=None, keepdims: bool = False) -> "DeviceArray":
        from . import device_ops as D
        return D.argmax(self, dim, keepdims)

    def transpose(self, *dims) -> "DeviceArray":
        from . import device_ops as D
        dims = dims[0] if len(dims) == 1 and isinstance(dims[0], (tuple, list)) else dims
        return D.permute(self, dims if dims else tuple(reversed(range(self.ndim))))

    def swapaxes(self, d1: int, d2: int) -> "DeviceArray":
        from . import device_ops as D
        return D.swapaxes(self, d1, d2)

    def squeeze(self) -> "DeviceArray":
        return self.reshape(tuple(d for d in self.shape if d != 1))

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


def current_workspace() -> Optional[DeviceArray]:
    """The scratch buffer in use right now (a CUDA graph captured against it keeps this reference, see graph.CapturedStep)."""
    return _ws


def _device_of(data: Any) -> Device:
    return cuda if isinstance(data, DeviceArray) else cpu


class Tensor:
    """Multi-dimensional array with a gradient slot (tensors.py:76-95)."""

    def __init__(self, data: Any) -> None:
        if isinstance(data, Tensor):
            data = data.data
        if not isinstance(data, (np.ndarray, DeviceArray)) and hasattr(data, "__cuda_array_interface__"):
            data = DeviceArray.from_cuda_array_interface(data)  # CuPy / Numba / torch device memory, adopted without a copy
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
        return Tensor(self.data.swapaxes(-1, -2))  # tensors.py:131-135: last two dims

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

    # ---- operators (tensors.py:176-306): `.data` is a NumPy array on cpu and a DeviceArray on cuda; both implement them
    def __getitem__(self, key: Any) -> "Tensor":
        return Tensor(self.data[_unwrap_key(key)])

    def __setitem__(self, key: Any, value) -> None:
        self.data[_unwrap_key(key)] = _unwrap(value)

    def __iter__(self) -> "Tensor":
        self._iterator = 0
        return self

    def __next__(self) -> "Tensor":
        if self._iterator == self.shape[0]:
            raise StopIteration
        y = self[self._iterator]
        self._iterator += 1
        return y

    def __add__(self, o): return Tensor(self.data + _unwrap(o))
    def __radd__(self, o): return Tensor(self.data + o) if o is not None else self.copy()  # tensors.py:199-201
    def __sub__(self, o): return Tensor(self.data - _unwrap(o))
    def __rsub__(self, o): return Tensor(o - self.data)
    def __mul__(self, o): return Tensor(self.data * _unwrap(o))
    def __rmul__(self, o): return Tensor(o * self.data)
    def __truediv__(self, o): return Tensor(self.data / _unwrap(o))
    def __rtruediv__(self, o): return Tensor(o / self.data)
    def __floordiv__(self, o): return Tensor(self.data // _unwrap(o))
    def __rfloordiv__(self, o): return Tensor(o // self.data)
    def __pow__(self, o): return Tensor(self.data ** _unwrap(o))
    def __rpow__(self, o): return Tensor(o ** self.data)
    def __mod__(self, o): return Tensor(self.data % _unwrap(o))
    def __rmod__(self, o): return Tensor(o % self.data)
    def __neg__(self): return Tensor(-self.data)
    def __invert__(self): return Tensor(~self.data)
    def __matmul__(self, o): return Tensor(self.data @ _unwrap(o))
    def __lt__(self, o): return Tensor(self.data < _unwrap(o))
    def __gt__(self, o): return Tensor(self.data > _unwrap(o))
    def __le__(self, o): return Tensor(self.data <= _unwrap(o))
    def __ge__(self, o): return Tensor(self.data >= _unwrap(o))
    def __eq__(self, o): return Tensor(self.data == _unwrap(o))  # type: ignore[override]
    def __ne__(self, o): return Tensor(self.data != _unwrap(o))  # type: ignore[override]

    def __hash__(self) -> int:
        return id(self)

    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        a = self.to_numpy()
        return a if dtype is None else a.astype(dtype)

    def _inplace(self, name: str, o) -> "Tensor":
        d = getattr(self.data, name)(_unwrap(o))  # NumPy / DeviceArray in-place operator: same array comes back
        if d is not NotImplemented:
            self.data = d
        return self

    def __iadd__(self, o): return self._inplace("__iadd__", o)
    def __isub__(self, o): return self._inplace("__isub__", o)
    def __imul__(self, o): return self._inplace("__imul__", o)
    def __itruediv__(self, o): return self._inplace("__itruediv__", o)
    def __ifloordiv__(self, o): return self._inplace("__ifloordiv__", o)
    def __ipow__(self, o): return self._inplace("__ipow__", o)
    def __imod__(self, o): return self._inplace("__imod__", o)

    # ---- dtype conversions (tensors.py:372-455)
    def to_type(self, dtype) -> "Tensor":
        dtype = np.dtype(dtype)
        return self if dtype == self.dtype else Tensor(self.data.astype(dtype))

    def ito_type(self, dtype) -> None:
        self.data = self.to_type(dtype).data

    def to_int(self): return self.to_type(np.int32)
    def to_long(self): return self.to_type(np.int64)
    def to_float(self): return self.to_type(np.float32)
    def to_list(self) -> list: return self.to_numpy().tolist()
    def to_cpu(self): return self.to_device(cpu)
    def to_cuda(self): return self.to_device(cuda)

    # ---- unary / reductions / shape (tensors.py:543-682); dim and keepdims as in the reference
    def abs(self): return Tensor(abs(self.data))
    def all(self, dim=None, *, keepdims=False): return Tensor(self.data.all(dim, keepdims=keepdims))
    def any(self, dim=None, *, keepdims=False): return Tensor(self.data.any(dim, keepdims=keepdims))
    def argmax(self, dim=None, *, keepdims=False): return Tensor(self.data.argmax(dim, keepdims=keepdims))
    def max(self, dim=None, *, keepdims=False): return Tensor(self.data.max(dim, keepdims=keepdims))
    def mean(self, dim=None, *, keepdims=False): return Tensor(self.data.mean(dim, keepdims=keepdims))
    def min(self, dim=None, *, keepdims=False): return Tensor(self.data.min(dim, keepdims=keepdims))
    def std(self, dim=None, *, keepdims=False): return Tensor(self.data.std(dim, keepdims=keepdims))
    def sum(self, dim=None, *, keepdims=False): return Tensor(self.data.sum(dim, keepdims=keepdims))
    def var(self, dim=None, *, ddof=0, keepdims=False): return Tensor(self.data.var(dim, ddof=ddof, keepdims=keepdims))
    def permute(self, dims): return Tensor(self.data.transpose(tuple(dims)))
    def transpose(self, dim1: int, dim2: int): return Tensor(self.data.swapaxes(dim1, dim2))
    def squeeze(self): return Tensor(self.data.squeeze())


def _unwrap(v: Any) -> Any:
    """``to_arraylike`` (tensors.py:685-689)"""
    return v.data if isinstance(v, Tensor) else v


def _unwrap_key(key: Any) -> Any:
    if isinstance(key, tuple):
        return tuple(_unwrap(k) for k in key)
    return _unwrap(key)


def tensor(data: Any, device: Optional[Device] = None, dtype=None) -> Tensor:
    """``compyute.tensor`` (tensors.py:44-73): host data → Tensor on ``device`` (default: context / cpu).  Device memory of
    another library (``__cuda_array_interface__``: CuPy, Numba, torch) becomes a cuda Tensor without a copy."""
    if not isinstance(data, (np.ndarray, Tensor)) and hasattr(data, "__cuda_array_interface__"):
        t = Tensor(DeviceArray.from_cuda_array_interface(data))
        if dtype is not None and np.dtype(dtype) != t.dtype:
            from . import device_ops as D
            t = Tensor(D.astype(t.data, dtype))
        return t
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
