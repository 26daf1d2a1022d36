"""NumPy-semantics operators on ``DeviceArray``.

The reference evaluates ``Tensor.data <op> other`` / ``Tensor.data.sum(dim)`` / ``device.module.<fn>(...)`` on CuPy
arrays (compyute/tensors.py:196-292, 552-682; compyute/tensor_ops/*.py).  Here the same calls land on the kernels of
``csrc/tensor_ops.cu`` through the C ABI; shapes, broadcasting and index arithmetic are resolved on the host.

Deliberate differences from CuPy/NumPy (documented, tested): arithmetic is carried in float32 (integer / bool operands
are cast; a float64 array is cast down), results of basic indexing / ``permute`` / ``transpose`` are copies unless the
selected region is one contiguous block (then a zero-copy view, like NumPy), boolean-mask and multi-array indexing raise
``NotImplementedError``.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes
import math
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib

MAX_DIMS = 6
DT = {"float32": 0, "int32": 1, "int64": 2, "bool": 3, "uint8": 3, "int8": 3, "float64": 4}
EW = {"add": 0, "sub": 1, "mul": 2, "div": 3, "pow": 4, "max": 5, "min": 6, "floordiv": 7, "mod": 8,
      "lt": 9, "gt": 10, "le": 11, "ge": 12, "eq": 13, "ne": 14}
UN = {"neg": 0, "abs": 1, "exp": 2, "log": 3, "log2": 4, "log10": 5, "sqrt": 6, "tanh": 7, "sin": 8, "cos": 9, "tan": 10,
      "sinh": 11, "cosh": 12, "clip": 13, "isnan": 14, "round": 15, "square": 16, "recip": 17}
RED = {"sum": 0, "sumsq": 1, "prod": 2, "max": 3, "min": 4, "any": 5, "all": 6, "count": 7, "argmax": 8}


def _T():
    from . import tensors
    return tensors


def _i64(seq: Sequence[int]):
    return (ctypes.c_int64 * max(len(seq), 1))(*[int(v) for v in seq])


def _contig_strides(shape: Sequence[int]) -> list[int]:
    s, acc = [], 1
    for d in reversed(shape):
        s.append(acc)
        acc *= int(d)
    return list(reversed(s))


def _bstrides(shape: Sequence[int], out_shape: Sequence[int]) -> list[int]:
    """Element strides of a C-contiguous array of ``shape`` walked over ``out_shape`` (0 on broadcast dims)."""
    pad = len(out_shape) - len(shape)
    st = _contig_strides(shape)
    return [0] * pad + [0 if d == 1 and o != 1 else s for d, s, o in zip(shape, st, out_shape[pad:])]


def _merge(dims, *strides):
    """Drops size-1 dims and merges neighbours that every operand walks contiguously, so the kernels' CPT_MAX_DIMS limit
    applies to the merged form (the C side does the same; this is only to validate the limit early)."""
    md: list[int] = []
    ms: list[list[int]] = [[] for _ in strides]
    for k, d in enumerate(dims):
        if d == 1:
            continue
        if md and all(s_[-1] == st[k] * d for s_, st in zip(ms, strides)):
            md[-1] *= d
            for s_, st in zip(ms, strides):
                s_[-1] = st[k]
        else:
            md.append(int(d))
            for s_, st in zip(ms, strides):
                s_.append(int(st[k]))
    return md, ms


def as_device(a: Any, dtype=None):
    """DeviceArray from a DeviceArray / Tensor / NumPy array / nested list (H2D copy for host data)."""
    t = _T()
    if isinstance(a, t.Tensor):
        a = a.data
    if isinstance(a, t.DeviceArray):
        return a if dtype is None or a.dtype == np.dtype(dtype) else astype(a, dtype)
    arr = np.asarray(a)
    if dtype is not None:
        arr = arr.astype(dtype, copy=False)
    elif arr.dtype == np.float64:
        arr = arr.astype(np.float32)
    return t.DeviceArray.from_numpy(arr)


def _f32(a):
    return a if a.dtype == np.float32 else astype(a, np.float32)


# ------------------------------------------------------------------------------------------------ elementwise
def astype(a, dtype):
    """tensors.py:372-455 — always a new array (like ``astype`` with a different dtype)."""
    t = _T()
    dtype = np.dtype(dtype)
    if dtype == a.dtype:
        return a.copy()
    if a.dtype.name not in DT or dtype.name not in DT:
        raise TypeError(f"astype: {a.dtype} -> {dtype} is not supported on the device")
    out = t.DeviceArray.empty(a.shape, dtype)
    _lib.check(_lib.lib().cpt_cast(out.ptr, DT[dtype.name], a.ptr, DT[a.dtype.name], a.size, t.stream_ptr()))
    return out


def binary(op: str, a, b, out=None, reverse: bool = False):
    """``a op b`` with NumPy broadcasting (``reverse``: ``b op a`` for a scalar b).  ``out`` (may be ``a``) makes it the
    in-place operator."""
    t = _T()
    code = EW[op]
    cmp_ = code >= EW["lt"]
    if isinstance(b, t.Tensor):
        b = b.data
    if isinstance(b, (list, tuple)) or (isinstance(b, np.ndarray) and b.ndim > 0):
        b = as_device(b)
    a32 = _f32(a)
    lib, st = _lib.lib(), t.stream_ptr()
    if not isinstance(b, t.DeviceArray):  # python / numpy scalar
        s = float(b)
        res = out if out is not None else t.DeviceArray.empty(a.shape, np.bool_ if cmp_ else np.float32)
        if out is not None and (out.shape != a.shape or out.dtype != np.float32 or cmp_):
            raise _T().ShapeError(f"in-place {op}: output {out.shape}/{out.dtype} does not match")
        dims = [a.size]
        _lib.check(lib.cpt_ew_binary(code, res.ptr, a32.ptr, None, s, 2 if reverse else 1, 1, _i64(dims), _i64([1]), _i64([0]), st))
        _drop_side_copies(res)
        return res
    if reverse:
        a32, b = _f32(b), a32
    b32 = _f32(b)
    try:
        oshape = tuple(int(v) for v in np.broadcast_shapes(a32.shape, b32.shape))
    except ValueError as e:
        raise t.ShapeError(str(e)) from None
    sa, sb = _bstrides(a32.shape, oshape), _bstrides(b32.shape, oshape)
    md, _ = _merge(oshape, sa, sb)
    if len(md) > MAX_DIMS:
        raise NotImplementedError(f"{op}: more than {MAX_DIMS} broadcast dims after merging")
    if out is not None:
        if out.shape != oshape or out.dtype != np.float32 or cmp_:
            raise t.ShapeError(f"in-place {op}: operand of shape {b32.shape} does not broadcast into {out.shape}")
        res = out
    else:
        res = t.DeviceArray.empty(oshape, np.bool_ if cmp_ else np.float32)
    nd = len(oshape)
    if nd > MAX_DIMS:  # pass the merged form
        oshape_k, (sa, sb) = md, _merge(oshape, sa, sb)[1]
        nd = len(oshape_k)
    else:
        oshape_k = oshape
    _lib.check(lib.cpt_ew_binary(code, res.ptr, a32.ptr, b32.ptr, 0.0, 0, nd, _i64(oshape_k), _i64(sa), _i64(sb), st))
    _drop_side_copies(res)
    return res


def _drop_side_copies(a) -> None:
    a.cl = a.stats = None


def unary(op: str, a, p0: float = 0.0, p1: float = 0.0, out=None):
    t = _T()
    a32 = _f32(a)
    res = out if out is not None else t.DeviceArray.empty(a.shape, np.bool_ if op == "isnan" else np.float32)
    _lib.check(_lib.lib().cpt_ew_unary(UN[op], res.ptr, a32.ptr, float(p0), float(p1), a.size, t.stream_ptr()))
    _drop_side_copies(res)
    return res


def logic(op: str, a, b=None):
    """bool arrays: and / or / xor / not (tensors.py:270)."""
    t = _T()
    code = {"and": 0, "or": 1, "xor": 2, "not": 3}[op]
    a8 = a if a.dtype == np.bool_ else astype(a, np.bool_)
    b8 = None
    if b is not None:
        b = as_device(b)
        b8 = b if b.dtype == np.bool_ else astype(b, np.bool_)
        if b8.shape != a8.shape:
            raise t.ShapeError(f"{op}: shapes {a8.shape} and {b8.shape} differ (no broadcasting of bool arrays on device)")
    out = t.DeviceArray.empty(a.shape, np.bool_)
    _lib.check(_lib.lib().cpt_logic(code, out.ptr, a8.ptr, b8.ptr if b8 is not None else None, a.size, t.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------ reductions
def _norm_dims(dim, ndim: int) -> tuple[int, ...]:
    if dim is None:
        return tuple(range(ndim))
    dims = (dim,) if isinstance(dim, (int, np.integer)) else tuple(dim)
    out = []
    for d in dims:
        d = int(d)
        if not -ndim <= d < ndim:
            raise _T().ShapeError(f"dim {d} out of range for a {ndim}-d array")
        out.append(d % ndim)
    if len(set(out)) != len(out):
        raise ValueError("duplicate value in 'dim'")
    return tuple(sorted(out))


def reduce(op: str, a, dim=None, keepdims: bool = False, scale: float = 1.0):
    t = _T()
    rdims = _norm_dims(dim, a.ndim)
    if a.dtype == np.bool_:
        x, xdt = a, DT["bool"]
    else:
        x, xdt = _f32(a), DT["float32"]
    kept_shape = tuple(1 if k in rdims else d for k, d in enumerate(a.shape)) if keepdims else \
        tuple(d for k, d in enumerate(a.shape) if k not in rdims)
    odt = np.int64 if op in ("count", "argmax") else np.bool_ if op in ("any", "all") else np.float32
    out = t.DeviceArray.empty(kept_shape, odt)
    n_red = int(np.prod([a.shape[k] for k in rdims], dtype=np.int64)) if rdims else 1
    if n_red == 0 and op in ("max", "min", "argmax"):
        raise ValueError(f"zero-size array to reduction operation {op} which has no identity")
    shape = list(a.shape)
    flags = [1 if k in rdims else 0 for k in range(a.ndim)]
    if len(shape) > MAX_DIMS:  # merge neighbours of the same kind
        ms, mf = [], []
        for d, f in zip(shape, flags):
            if ms and mf[-1] == f:
                ms[-1] *= d
            else:
                ms.append(d); mf.append(f)
        shape, flags = ms, mf
        if len(shape) > MAX_DIMS:
            raise NotImplementedError("reduce: too many alternating kept / reduced axes")
    lib = _lib.lib()
    ws, ws_bytes = t.workspace(lib.cpt_reduce_workspace_size(max(out.size, 1)))
    _lib.check(lib.cpt_reduce(RED[op], out.ptr, x.ptr, xdt, len(shape), _i64(shape), (ctypes.c_int32 * max(len(shape), 1))(*flags),
                              float(scale), ws, ws_bytes, t.stream_ptr()))
    return out


def sum_(a, dim=None, keepdims=False):
    return reduce("count" if a.dtype == np.bool_ else "sum", a, dim, keepdims)


def _count(a, dim) -> int:
    return int(np.prod([a.shape[k] for k in _norm_dims(dim, a.ndim)], dtype=np.int64)) if a.ndim else 1


def mean(a, dim=None, keepdims=False):
    n = _count(a, dim)
    return reduce("sum", _f32(a), dim, keepdims, scale=(1.0 / n) if n else float("nan"))


def var(a, dim=None, ddof: int = 0, keepdims=False):
    """NumPy's own two-pass algorithm (mean, then the mean of squared deviations)."""
    n = _count(a, dim)
    a32 = _f32(a)
    mu = mean(a32, dim, keepdims=True)
    dev = binary("sub", a32, mu)
    unary("square", dev, out=dev)
    return reduce("sum", dev, dim, keepdims, scale=1.0 / (n - ddof) if n - ddof > 0 else float("nan"))


def std(a, dim=None, keepdims=False):
    v = var(a, dim, 0, keepdims)
    return unary("sqrt", v, out=v)


def norm(a, dim=None, keepdims=False):
    """reduction_ops.py:90-109 (``linalg.norm``: Frobenius / 2-norm over the given dims)."""
    v = reduce("sumsq", _f32(a), dim, keepdims)
    return unary("sqrt", v, out=v)


def argmax(a, dim=None, keepdims=False):
    if dim is not None and not isinstance(dim, (int, np.integer)):
        raise TypeError("argmax: dim must be an int or None")
    out = reduce("argmax", _f32(a), dim, keepdims=False)
    if keepdims:
        shape = tuple(1 for _ in a.shape) if dim is None else tuple(1 if k == int(dim) % a.ndim else d for k, d in enumerate(a.shape))
        out = out.reshape(shape)
    return out


# ------------------------------------------------------------------------------------------------ data movement
def _copy_strided(dst, dst_off: int, src, src_off: int, dims, dst_strides, src_strides) -> None:
    t = _T()
    item = dst.dtype.itemsize
    if item not in (1, 4, 8) or src.dtype.itemsize != item:
        raise TypeError(f"strided copy: element sizes {src.dtype} -> {dst.dtype}")
    md, (mds, mss) = _merge(dims, dst_strides, src_strides)
    if len(md) > MAX_DIMS:
        raise NotImplementedError(f"more than {MAX_DIMS} dims after merging")
    _lib.check(_lib.lib().cpt_strided_copy(dst.ptr + dst_off * item, src.ptr + src_off * item, item, len(md), _i64(md), _i64(mds),
                                           _i64(mss), t.stream_ptr()))
    _drop_side_copies(dst)


def _basic_index(shape, key):
    """(offset, dims, strides) in elements of the region a basic index (ints, slices, Ellipsis, None) selects in a
    C-contiguous array of ``shape``."""
    if not isinstance(key, tuple):
        key = (key,)
    n_spec = sum(1 for k in key if k is not None and k is not Ellipsis)
    if sum(1 for k in key if k is Ellipsis) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    if n_spec > len(shape):
        raise IndexError(f"too many indices for array: array is {len(shape)}-dimensional, but {n_spec} were indexed")
    e = next((i for i, k in enumerate(key) if k is Ellipsis), None)
    if e is not None:
        key = key[:e] + (slice(None),) * (len(shape) - n_spec) + key[e + 1:]
    else:
        key = key + (slice(None),) * (len(shape) - n_spec)
    st = _contig_strides(shape)
    off, dims, strides, ax = 0, [], [], 0
    for k in key:
        if k is None:
            dims.append(1); strides.append(0)
            continue
        d = shape[ax]
        if isinstance(k, (int, np.integer)):
            i = int(k)
            if not -d <= i < d:
                raise IndexError(f"index {i} is out of bounds for axis {ax} with size {d}")
            off += (i % d) * st[ax]
        elif isinstance(k, slice):
            start, stop, step = k.indices(d)
            n = max(0, (stop - start + (step - (1 if step > 0 else -1))) // step)
            off += start * st[ax] if n else 0
            dims.append(n); strides.append(st[ax] * step)
        else:
            raise NotImplementedError(f"index of type {type(k).__name__} inside a tuple index")
        ax += 1
    return off, dims, strides


def _is_index_array(key) -> bool:
    t = _T()
    if isinstance(key, (t.Tensor, t.DeviceArray)):
        return True
    if isinstance(key, np.ndarray):
        return key.ndim > 0
    return isinstance(key, list)


def getitem(a, key):
    """tensors.py:176-180.  Basic indexing → zero-copy view when the region is one contiguous block, else a copy;
    an integer index array on the leading axis → row gather."""
    t = _T()
    if _is_index_array(key):
        idx = as_device(key)
        if idx.dtype == np.bool_:
            raise NotImplementedError("boolean-mask indexing is not implemented on the device")
        if idx.dtype not in (np.int32, np.int64):
            raise IndexError("arrays used as indices must be of integer type")
        if a.ndim == 0:
            raise IndexError("too many indices for array")
        row = int(np.prod(a.shape[1:], dtype=np.int64))
        out = t.DeviceArray.empty(tuple(idx.shape) + tuple(a.shape[1:]), a.dtype)
        if out.size:
            _lib.check(_lib.lib().cpt_gather_rows(out.ptr, a.ptr, idx.ptr, DT[idx.dtype.name], idx.size, row * a.dtype.itemsize, a.shape[0],
                                                  None, t.stream_ptr()))
        return out
    off, dims, strides = _basic_index(a.shape, key)
    n = int(np.prod(dims, dtype=np.int64)) if dims else 1
    if strides == _contig_strides(dims) or n == 0 or all(d == 1 or s == c for d, s, c in zip(dims, strides, _contig_strides(dims))):
        return t.DeviceArray(a._buf.view(-1)[off:off + n], tuple(dims), a.dtype)  # contiguous block: a view
    out = t.DeviceArray.empty(tuple(dims), a.dtype)
    _copy_strided(out, 0, a, off, dims, _contig_strides(dims), strides)
    return out


def setitem(a, key, value) -> None:
    """tensors.py:182-183 for basic indices: ``a[key] = value`` (scalar or broadcastable array)."""
    t = _T()
    if _is_index_array(key):
        raise NotImplementedError("assignment through an index array is not implemented on the device")
    off, dims, strides = _basic_index(a.shape, key)
    if isinstance(value, t.Tensor):
        value = value.data
    if not isinstance(value, t.DeviceArray):
        value = t.DeviceArray.from_numpy(np.asarray(value, dtype=a.dtype))
    elif value.dtype != a.dtype:
        value = astype(value, a.dtype)
    try:
        if tuple(np.broadcast_shapes(tuple(dims), value.shape)) != tuple(dims):
            raise ValueError
    except ValueError:
        raise t.ShapeError(f"could not broadcast input array from shape {value.shape} into shape {tuple(dims)}") from None
    _copy_strided(a, off, value, 0, dims, strides, _bstrides(value.shape, dims))


def permute(a, dims: Sequence[int]):
    """tensors.py:615-622 / shape_ops.py:230 — materialised (a DeviceArray is always C-contiguous)."""
    t = _T()
    dims = tuple(int(d) % a.ndim for d in dims)
    if sorted(dims) != list(range(a.ndim)):
        raise ValueError("axes don't match array")
    st = _contig_strides(a.shape)
    oshape = tuple(a.shape[d] for d in dims)
    if dims == tuple(range(a.ndim)):
        return a
    out = t.DeviceArray.empty(oshape, a.dtype)
    _copy_strided(out, 0, a, 0, oshape, _contig_strides(oshape), [st[d] for d in dims])
    return out


def swapaxes(a, d1: int, d2: int):
    p = list(range(a.ndim))
    p[d1], p[d2] = p[d2], p[d1]
    return permute(a, p)


def flip(a, dim=None):
    """shape_ops.py:123-140"""
    t = _T()
    fd = _norm_dims(dim, a.ndim)
    st = _contig_strides(a.shape)
    off = sum((a.shape[k] - 1) * st[k] for k in fd) if a.size else 0
    out = t.DeviceArray.empty(a.shape, a.dtype)
    _copy_strided(out, 0, a, off, a.shape, st, [-s if k in fd else s for k, s in enumerate(st)])
    return out


def broadcast_to(a, shape):
    t = _T()
    shape = tuple(int(s) for s in shape)
    if tuple(np.broadcast_shapes(a.shape, shape)) != shape:
        raise t.ShapeError(f"cannot broadcast {a.shape} to {shape}")
    out = t.DeviceArray.empty(shape, a.dtype)
    _copy_strided(out, 0, a, 0, shape, _contig_strides(shape), _bstrides(a.shape, shape))
    return out


def full(shape, value, dtype=np.float32):
    t = _T()
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    dtype = np.dtype(dtype)
    out = t.DeviceArray.empty(shape, dtype)
    if out.size == 0:
        return out
    if dtype == np.float32:
        out.fill(float(value))
    elif value == 0:
        out._buf.zero_()  # cudaMemsetAsync
    else:
        one = full((1,), float(value))
        src = one if dtype == np.float32 else astype(one, dtype)
        _copy_strided(out, 0, src, 0, [out.size], [1], [0])
    return out


def pad(a, widths):
    """shape_ops.py:186-206: zero padding, ``widths`` = ((before, after), ...) per dim."""
    widths = [tuple(int(v) for v in w) for w in widths]
    if len(widths) != a.ndim:
        raise _T().ShapeError("pad: one (before, after) pair per dimension")
    out = full(tuple(d + w[0] + w[1] for d, w in zip(a.shape, widths)), 0, a.dtype)
    setitem(out, tuple(slice(w[0], w[0] + d) for d, w in zip(a.shape, widths)), a)
    return out


def concat(arrays, dim: int = -1):
    """shape_ops.py:72-88"""
    t = _T()
    arrays = [as_device(x) for x in arrays]
    nd = arrays[0].ndim
    dim = int(dim) % nd
    for x in arrays[1:]:
        if x.ndim != nd or any(x.shape[k] != arrays[0].shape[k] for k in range(nd) if k != dim) or x.dtype != arrays[0].dtype:
            raise t.ShapeError("concat: all input arrays must match in dtype and in every dim but the concatenation dim")
    oshape = list(arrays[0].shape)
    oshape[dim] = sum(x.shape[dim] for x in arrays)
    out = t.DeviceArray.empty(tuple(oshape), arrays[0].dtype)
    pos = 0
    for x in arrays:
        key = tuple(slice(pos, pos + x.shape[dim]) if k == dim else slice(None) for k in range(nd))
        setitem(out, key, x)
        pos += x.shape[dim]
    return out


def stack(arrays, dim: int = 0):
    arrays = [as_device(x) for x in arrays]
    dim = int(dim) % (arrays[0].ndim + 1)
    return concat([x.reshape(x.shape[:dim] + (1,) + x.shape[dim:]) for x in arrays], dim)


def split(a, splits, dim: int = -1):
    """shape_ops.py:380-399 (``numpy.split``: number of equal parts or the split positions)."""
    dim = int(dim) % a.ndim
    n = a.shape[dim]
    if isinstance(splits, (int, np.integer)):
        if n % int(splits):
            raise ValueError("array split does not result in an equal division")
        step = n // int(splits)
        edges = list(range(0, n + 1, step))
    else:
        edges = [0] + [int(s) for s in splits] + [n]
    return [getitem(a, tuple(slice(lo, max(lo, hi)) if k == dim else slice(None) for k in range(a.ndim)))
            for lo, hi in zip(edges[:-1], edges[1:])]


def tile(a, n_repeats: int, dim: int):
    """shape_ops.py:436-455"""
    return concat([a] * int(n_repeats), dim)


def repeat(a, n: int, dim: int):
    """shape_ops.py:308-359 building block: every element repeated n times along ``dim``."""
    dim = int(dim) % a.ndim
    x = a.reshape(a.shape[:dim + 1] + (1,) + a.shape[dim + 1:])
    x = broadcast_to(x, a.shape[:dim + 1] + (int(n),) + a.shape[dim + 1:])
    return x.reshape(a.shape[:dim] + (a.shape[dim] * int(n),) + a.shape[dim + 1:])


def arange(stop, start=0, step=1, dtype=np.int64):
    """creation_ops.py:24-57"""
    t = _T()
    dtype = np.dtype(dtype)
    n = max(0, int(math.ceil((stop - start) / step)))
    out = t.DeviceArray.empty((n,), dtype)
    _lib.check(_lib.lib().cpt_arange(out.ptr, DT[dtype.name], float(start), float(step), n, t.stream_ptr()))
    return out


def identity(n: int, dtype=np.float32):
    out = full((n, n), 0, dtype)
    if n:
        _copy_strided(out, 0, full((1,), 1, dtype), 0, [n], [n + 1], [0])
    return out


_seed_state = {"seed": 0x5EED, "calls": 0}


def set_seed(value: Optional[int]) -> None:
    """random/random.py:25-36"""
    _seed_state["seed"] = 0x5EED if value is None else int(value)
    _seed_state["calls"] = 0


def random_fill(shape, kind: int, p0: float, p1: float):
    t = _T()
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    out = t.DeviceArray.empty(shape, np.float32)
    _seed_state["calls"] += 1
    seed = (_seed_state["seed"] * 0x9E3779B97F4A7C15 + _seed_state["calls"] * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
    _lib.check(_lib.lib().cpt_random_fill(out.ptr, out.size, kind, float(p0), float(p1), seed, t.stream_ptr()))
    return out


def matmul(a, b):
    """tensors.py:273-274 for (..., N, K) @ (K, M) and batched (..., N, K) @ (..., K, M): the Linear dgrad contraction
    ``dy @ w`` of the C ABI (linear_funcs.py:31), in the current compute mode."""
    t = _T()
    from .backend import get_compute_mode
    a, b = _f32(as_device(a)), _f32(as_device(b))
    if a.ndim < 1 or b.ndim < 1:
        raise ValueError("matmul: input operand does not have enough dimensions")
    if a.ndim == 1 or b.ndim == 1:
        a2 = a.reshape(1, -1) if a.ndim == 1 else a
        b2 = b.reshape(-1, 1) if b.ndim == 1 else b
        y = matmul(a2, b2)
        shape = list(y.shape)
        if b.ndim == 1:
            shape.pop(-1)
        if a.ndim == 1:
            shape.pop(-1 if b.ndim == 1 else -2)
        return y.reshape(tuple(shape))
    if a.shape[-1] != b.shape[-2]:
        raise t.ShapeError(f"matmul: shapes {a.shape} and {b.shape} not aligned")
    lib, mode = _lib.lib(), get_compute_mode()
    K, M = b.shape[-2], b.shape[-1]

    def gemm(x, w, y, n):
        ws, wsb = t.workspace(lib.cpt_linear_workspace_size(_lib.OP_DGRAD, n, M, K, mode))
        _lib.check(lib.cpt_linear_dgrad(x.ptr, w.ptr, y.ptr, n, M, K, mode, ws, wsb, t.stream_ptr()))

    if b.ndim == 2:
        n = a.size // K
        out = t.DeviceArray.empty(a.shape[:-1] + (M,), np.float32)
        if out.size:
            gemm(a, b, out, n)
        return out
    batch = tuple(int(v) for v in np.broadcast_shapes(a.shape[:-2], b.shape[:-2]))
    N = a.shape[-2]
    ab = broadcast_to(a, batch + a.shape[-2:]) if a.shape[:-2] != batch else a
    bb = broadcast_to(b, batch + b.shape[-2:]) if b.shape[:-2] != batch else b
    out = t.DeviceArray.empty(batch + (N, M), np.float32)
    nb = int(np.prod(batch, dtype=np.int64))
    a3, b3, o3 = ab.reshape(nb, N, K), bb.reshape(nb, K, M), out.reshape(nb, N, M)
    for i in range(nb):
        gemm(getitem(a3, i), getitem(b3, i), getitem(o3, i), N)
    return out


def allclose(a, b, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    """multiary_ops.py:17-36: all(|a - b| <= atol + rtol·|b|)"""
    a, b = as_device(a), as_device(b)
    lhs = unary("abs", binary("sub", a, b))
    rhs = binary("add", binary("mul", unary("abs", b), rtol), atol)
    return bool(reduce("all", binary("le", lhs, rhs)).item())
