"""Function forms of the device-tensor operators; names / arguments follow compyute/tensor_ops/*.py."""

from __future__ import annotations

import builtins
from typing import Optional, Sequence

import numpy as np

from .. import device_ops as D
from ..backend import Device, select_device
from ..tensors import DeviceArray, ShapeError, Tensor

__all__ = [
    # creation_ops.py
    "arange", "empty", "empty_like", "full", "full_like", "identity", "linspace", "ones", "ones_like", "zeros", "zeros_like",
    # unary_ops.py
    "abs", "clip", "cos", "cosh", "exp", "is_nan", "log", "log2", "log10", "round", "sech", "sin", "sinh", "sqrt", "tan", "tanh",
    # reduction_ops.py
    "all", "any", "mean", "norm", "prod", "std", "sum", "tensorsum", "tensorprod", "var",
    # selection_ops.py
    "argmax", "max", "maximum", "min", "minimum",
    # shape_ops.py
    "append", "broadcast_to", "concat", "flatten", "flip", "insert_dim", "movedim", "pad", "pad_to_shape", "permute", "repeat1d",
    "repeat2d", "reshape", "split", "squeeze", "stack", "tile", "transpose",
    # multiary_ops.py
    "allclose", "dot", "inner", "outer",
]


def _on_device(x: Tensor) -> bool:
    return isinstance(x.data, DeviceArray)


def _shape(shape) -> tuple[int, ...]:
    return tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))


# ------------------------------------------------------------------------------------------------ creation_ops.py
def _create(shape, value, dtype, device: Optional[Device]) -> Tensor:
    dtype = np.dtype(dtype or np.float32)
    if select_device(device).t == "cuda":
        return Tensor(D.full(_shape(shape), value, dtype))
    return Tensor(np.full(_shape(shape), value, dtype))


def empty(shape, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:60-86"""
    if select_device(device).t == "cuda":
        return Tensor(DeviceArray.empty(_shape(shape), np.dtype(dtype or np.float32)))
    return Tensor(np.empty(_shape(shape), np.dtype(dtype or np.float32)))


def zeros(shape, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:265-291"""
    return _create(shape, 0, dtype, device)


def ones(shape, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:220-246"""
    return _create(shape, 1, dtype, device)


def full(shape, value, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:105-134"""
    return _create(shape, value, dtype, device)


def empty_like(x: Tensor) -> Tensor:
    return empty(x.shape, device=x.device, dtype=x.dtype)


def zeros_like(x: Tensor) -> Tensor:
    return zeros(x.shape, device=x.device, dtype=x.dtype)


def ones_like(x: Tensor) -> Tensor:
    return ones(x.shape, device=x.device, dtype=x.dtype)


def full_like(x: Tensor, value) -> Tensor:
    return full(x.shape, value, device=x.device, dtype=x.dtype)


def arange(stop, start=0, step=1, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:24-57"""
    dtype = np.dtype(dtype or np.int64)
    if select_device(device).t == "cuda":
        return Tensor(D.arange(stop, start, step, dtype))
    return Tensor(np.arange(start, stop, step, dtype))


def identity(n: int, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:155-181"""
    dtype = np.dtype(dtype or np.float32)
    if select_device(device).t == "cuda":
        return Tensor(D.identity(int(n), dtype))
    return Tensor(np.identity(int(n), dtype))


def linspace(start: float, stop: float, num: int, *, device: Optional[Device] = None, dtype=None) -> Tensor:
    """creation_ops.py:184-217"""
    dtype = np.dtype(dtype or np.float32)
    if select_device(device).t == "cuda":
        step = (stop - start) / (num - 1) if num > 1 else 0.0
        a = D.arange(num, 0, 1, np.float32)
        a = D.binary("add", D.binary("mul", a, step), start)
        return Tensor(a if dtype == np.float32 else D.astype(a, dtype))
    return Tensor(np.linspace(start, stop, num, dtype=dtype))


# ------------------------------------------------------------------------------------------------ unary_ops.py
def _unary(name: str, npf, x: Tensor, *p) -> Tensor:
    return Tensor(D.unary(name, x.data, *p)) if _on_device(x) else Tensor(npf(x.data))


def abs(x: Tensor) -> Tensor: return _unary("abs", np.abs, x)  # noqa: A001
def cos(x: Tensor) -> Tensor: return _unary("cos", np.cos, x)
def cosh(x: Tensor) -> Tensor: return _unary("cosh", np.cosh, x)
def exp(x: Tensor) -> Tensor: return _unary("exp", np.exp, x)
def is_nan(x: Tensor) -> Tensor: return _unary("isnan", np.isnan, x)
def log(x: Tensor) -> Tensor: return _unary("log", np.log, x)
def log2(x: Tensor) -> Tensor: return _unary("log2", np.log2, x)
def log10(x: Tensor) -> Tensor: return _unary("log10", np.log10, x)
def sin(x: Tensor) -> Tensor: return _unary("sin", np.sin, x)
def sinh(x: Tensor) -> Tensor: return _unary("sinh", np.sinh, x)
def sqrt(x: Tensor) -> Tensor: return _unary("sqrt", np.sqrt, x)
def tan(x: Tensor) -> Tensor: return _unary("tan", np.tan, x)
def tanh(x: Tensor) -> Tensor: return _unary("tanh", np.tanh, x)


def sech(x: Tensor) -> Tensor:
    """unary_ops.py:358-371: 1 / cosh(x)"""
    return Tensor(D.unary("recip", D.unary("cosh", x.data))) if _on_device(x) else Tensor(1 / np.cosh(x.data))


def clip(x: Tensor, min_value: Optional[float] = None, max_value: Optional[float] = None) -> Tensor:
    """unary_ops.py:49-68"""
    if _on_device(x):
        lo = -np.inf if min_value is None else min_value
        hi = np.inf if max_value is None else max_value
        return Tensor(D.unary("clip", x.data, lo, hi))
    return Tensor(np.clip(x.data, min_value, max_value))


def round(x: Tensor, decimals: int) -> Tensor:  # noqa: A001
    """unary_ops.py:340-355"""
    return Tensor(D.unary("round", x.data, 10.0 ** int(decimals))) if _on_device(x) else Tensor(np.round(x.data, decimals))


# ------------------------------------------------------------------------------------------------ reduction_ops.py / selection_ops.py
def all(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.all(dim, keepdims=keepdims)  # noqa: A001
def any(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.any(dim, keepdims=keepdims)  # noqa: A001
def mean(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.mean(dim, keepdims=keepdims)
def std(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.std(dim, keepdims=keepdims)
def sum(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.sum(dim, keepdims=keepdims)  # noqa: A001
def var(x: Tensor, dim=None, *, ddof: int = 0, keepdims: bool = False) -> Tensor: return x.var(dim, ddof=ddof, keepdims=keepdims)
def argmax(x: Tensor, dim: Optional[int] = None, *, keepdims: bool = False) -> Tensor: return x.argmax(dim, keepdims=keepdims)
def max(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.max(dim, keepdims=keepdims)  # noqa: A001
def min(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor: return x.min(dim, keepdims=keepdims)  # noqa: A001


def prod(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor:
    """reduction_ops.py:112-131"""
    return Tensor(x.data.prod(dim, keepdims=keepdims))


def norm(x: Tensor, dim=None, *, keepdims: bool = False) -> Tensor:
    """reduction_ops.py:90-109"""
    if _on_device(x):
        return Tensor(D.norm(x.data, dim, keepdims))
    return Tensor(np.asarray(np.linalg.norm(x.data, axis=dim, keepdims=keepdims)))


def tensorsum(tensors) -> Tensor:
    """reduction_ops.py:194-207"""
    it = iter(tensors)
    acc = next(it).copy()
    for t in it:
        acc += t
    return acc


def tensorprod(tensors) -> Tensor:
    """reduction_ops.py:178-191"""
    it = iter(tensors)
    acc = next(it).copy()
    for t in it:
        acc *= t
    return acc


def maximum(x1: Tensor, x2) -> Tensor:
    """selection_ops.py:89-104"""
    o = x2.data if isinstance(x2, Tensor) else x2
    return Tensor(D.binary("max", x1.data, o)) if _on_device(x1) else Tensor(np.maximum(x1.data, o))


def minimum(x1: Tensor, x2) -> Tensor:
    """selection_ops.py:129-144"""
    o = x2.data if isinstance(x2, Tensor) else x2
    return Tensor(D.binary("min", x1.data, o)) if _on_device(x1) else Tensor(np.minimum(x1.data, o))


# ------------------------------------------------------------------------------------------------ shape_ops.py
def reshape(x: Tensor, shape) -> Tensor:
    return Tensor(x.data.reshape(_shape(shape)))


def flatten(x: Tensor) -> Tensor:
    return Tensor(x.data.reshape((-1,)))


def squeeze(x: Tensor) -> Tensor:
    return x.squeeze()


def insert_dim(x: Tensor, dim: int) -> Tensor:
    """shape_ops.py:143-162"""
    nd = x.ndim + 1
    dim = dim % nd
    return Tensor(x.data.reshape(x.shape[:dim] + (1,) + x.shape[dim:]))


def transpose(x: Tensor, dim1: int, dim2: int) -> Tensor:
    return x.transpose(dim1, dim2)


def permute(x: Tensor, dims) -> Tensor:
    return x.permute(tuple(dims))


def movedim(x: Tensor, from_dim: int, to_dim: int) -> Tensor:
    """shape_ops.py:165-183"""
    order = [d for d in range(x.ndim) if d != from_dim % x.ndim]
    order.insert(to_dim % x.ndim, from_dim % x.ndim)
    return x.permute(tuple(order))


def flip(x: Tensor, dim=None) -> Tensor:
    return Tensor(D.flip(x.data, dim)) if _on_device(x) else Tensor(np.flip(x.data, dim))


def broadcast_to(x: Tensor, shape) -> Tensor:
    return Tensor(D.broadcast_to(x.data, _shape(shape))) if _on_device(x) else Tensor(np.broadcast_to(x.data, _shape(shape)))


def concat(tensors: Sequence[Tensor], dim: int = -1) -> Tensor:
    if _on_device(tensors[0]):
        return Tensor(D.concat([t.data for t in tensors], dim))
    return Tensor(np.concatenate([t.data for t in tensors], axis=dim))


def append(x: Tensor, values: Tensor, dim: int = -1) -> Tensor:
    return concat([x, values], dim)


def stack(tensors: Sequence[Tensor], dim: int = 0) -> Tensor:
    if _on_device(tensors[0]):
        return Tensor(D.stack([t.data for t in tensors], dim))
    return Tensor(np.stack([t.data for t in tensors], axis=dim))


def split(x: Tensor, splits, dim: int = -1) -> list[Tensor]:
    if _on_device(x):
        return [Tensor(a) for a in D.split(x.data, splits, dim)]
    return [Tensor(a) for a in np.split(x.data, splits, axis=dim)]


def tile(x: Tensor, n_repeats: int, dim: int) -> Tensor:
    if _on_device(x):
        return Tensor(D.tile(x.data, n_repeats, dim))
    reps = [1] * x.ndim
    reps[dim] = n_repeats
    return Tensor(np.tile(x.data, tuple(reps)))


def repeat1d(x: Tensor, n: int) -> Tensor:
    """shape_ops.py:308-332: repeat along the last dim"""
    return Tensor(D.repeat(x.data, n, -1)) if _on_device(x) else Tensor(np.repeat(x.data, n, axis=-1))


def repeat2d(x: Tensor, n: int) -> Tensor:
    """shape_ops.py:335-359: repeat along the last two dims (upsampling)"""
    if _on_device(x):
        return Tensor(D.repeat(D.repeat(x.data, n, -1), n, -2))
    return Tensor(np.repeat(np.repeat(x.data, n, axis=-1), n, axis=-2))


def pad(x: Tensor, padding) -> Tensor:
    """shape_ops.py:186-206: int, (before, after) or one (before, after) per dim"""
    if isinstance(padding, (int, np.integer)):
        widths = [(int(padding), int(padding))] * x.ndim
    elif len(padding) == 2 and builtins.all(isinstance(p, (int, np.integer)) for p in padding):
        widths = [tuple(padding)] * x.ndim
    else:
        widths = [tuple(p) for p in padding]
    return Tensor(D.pad(x.data, widths)) if _on_device(x) else Tensor(np.pad(x.data, widths))


def pad_to_shape(x: Tensor, shape) -> Tensor:
    """shape_ops.py:209-227: zero padding at the end of each dim"""
    if len(shape) != x.ndim or builtins.any(s < d for s, d in zip(shape, x.shape)):
        raise ShapeError(f"cannot pad {x.shape} to {tuple(shape)}")
    return pad(x, [(0, int(s) - d) for s, d in zip(shape, x.shape)])


# ------------------------------------------------------------------------------------------------ multiary_ops.py
def allclose(x1: Tensor, x2: Tensor, rtol: float = 1e-05, atol: float = 1e-08) -> bool:
    if _on_device(x1) or _on_device(x2):
        return D.allclose(x1.data, x2.data, rtol, atol)
    return bool(np.allclose(x1.data, x2.data, rtol, atol))


def dot(x1: Tensor, x2: Tensor) -> Tensor:
    """multiary_ops.py:76-91"""
    return x1 @ x2 if (x1.ndim > 1 or x2.ndim > 1) else inner(x1, x2)


def inner(*tensors: Tensor) -> Tensor:
    """multiary_ops.py:115-128 for two vectors"""
    if len(tensors) != 2:
        raise NotImplementedError("inner: two operands")
    a, b = tensors
    return (a * b).sum(-1)


def outer(*tensors: Tensor) -> Tensor:
    """multiary_ops.py:131-144 for two vectors"""
    if len(tensors) != 2:
        raise NotImplementedError("outer: two operands")
    a, b = tensors
    return Tensor(a.data.reshape((-1, 1))) * Tensor(b.data.reshape((1, -1)))
