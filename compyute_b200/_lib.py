"""ctypes binding of libcompyute_b200.so — the C-ABI boundary (include/compyute_b200.h).

Prototypes are generated from the header itself so the Python side can never drift from the
declared ABI.  There is NO fallback: if the library is missing, ``lib()`` raises.
"""

from __future__ import annotations

import ctypes
import os
import re
from functools import lru_cache

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "compyute_b200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libcompyute_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, -1, -2, -3, -4
MODE_FP32, MODE_TF32, MODE_BF16, MODE_FP32X3 = 0, 1, 2, 3
OP_FPROP, OP_DGRAD, OP_WGRAD = 0, 1, 2


class CudaRuntimeError(RuntimeError):
    """Mirrors ``compyute.backend.CUDARuntimeError`` (backend.py:18-22)."""


class ConvDesc(ctypes.Structure):
    """``cpt_conv2d_desc``"""

    _fields_ = [(n, ctypes.c_int32) for n in ("B", "Ci", "H", "W", "Co", "K", "pad", "stride", "dil")]


class ParamEntry(ctypes.Structure):
    """``cpt_param_entry``"""

    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("n", ctypes.c_int64)]


class DpView(ctypes.Structure):
    """``cpt_dp_view``"""

    _fields_ = [("p_local", ctypes.c_void_p), ("p_mc", ctypes.c_void_p), ("p_peers", ctypes.c_void_p), ("g_local", ctypes.c_void_p),
                ("g_mc", ctypes.c_void_p), ("g_peers", ctypes.c_void_p), ("shard_off", ctypes.c_int64), ("shard_elems", ctypes.c_int64),
                ("world", ctypes.c_int32), ("pre_reduced", ctypes.c_int32), ("max_ctas_per_sm", ctypes.c_int32), ("_pad", ctypes.c_int32)]


_SCALARS = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
            "float": ctypes.c_float, "double": ctypes.c_double, "uint64_t": ctypes.c_uint64, "void": None}


def parse_header(path: str = HEADER) -> dict[str, tuple[object, list[object]]]:
    """Returns {symbol: (restype, [argtypes])} for every ``cpt_*`` function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^[ \t]*#[^\n]*$", " ", src, flags=re.M)  # preprocessor lines
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w\s\*]*?)\b(cpt_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret = ret.strip()
        if "typedef" in ret or "struct" in ret:
            continue
        restype = ctypes.c_char_p if "char" in ret else _SCALARS[ret.replace("const", "").strip()]
        argtypes = []
        for a in [s.strip() for s in args.split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                argtypes.append(_SCALARS[a.replace("const", "").split()[0]])
        protos[name] = (restype, argtypes)
    return protos


@lru_cache(maxsize=1)
def lib() -> ctypes.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C compyute_b200/csrc`). compyute_b200 has no CPU fallback.")
    dll = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(dll, name)  # AttributeError here == header/library drift
        fn.restype = restype
        fn.argtypes = argtypes
    return dll


def check(code: int) -> None:
    """Maps a negative return code onto the exception the reference would raise."""
    if code == OK:
        return
    msg = lib().cpt_last_error().decode()
    if code == ERR_INVALID:
        from .tensors import ShapeError
        raise ShapeError(msg)
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if code == ERR_WORKSPACE:
        raise ValueError(msg)
    raise CudaRuntimeError(msg)
