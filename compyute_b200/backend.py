"""Devices and compute mode.  Mirrors compyute/backend.py:26-178 for the in-scope surface.

``cpu`` tensors hold NumPy arrays and exist only as host staging (creation, ``to_numpy``); every hot-path
Function requires ``cuda`` tensors and raises ``DeviceError`` otherwise — there is no CPU fallback.
Unlike the reference, ``import cupy`` is not required (CuPy is not in this image): device memory is owned by
torch's caching allocator and handed to the C ABI as raw pointers.
"""

from __future__ import annotations

import os
from contextlib import contextmanager
from dataclasses import dataclass

from . import _lib

__all__ = ["Device", "cpu", "cuda", "DeviceError", "gpu_available", "synchronize", "use_device", "select_device",
           "compute_mode", "get_compute_mode", "set_compute_mode", "set_strip_conv_enabled"]


class DeviceError(Exception):
    """Tensors on mismatching / unsupported devices (backend.py:15)."""


@dataclass(frozen=True, repr=False)
class Device:
    t: str
    index: int = 0

    def __repr__(self) -> str:
        return f'Device("{self.t}:{self.index}")' if self.t == "cuda" else 'Device("cpu")'

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


cpu = Device("cpu")


def _local_cuda_index() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


class _Cuda(Device):
    """The process's CUDA device (one process per GPU: index = LOCAL_RANK)."""

    def __init__(self):
        super().__init__("cuda", _local_cuda_index())


cuda = _Cuda()


def gpu_available() -> bool:
    """backend.py:102-114"""
    import torch
    return torch.cuda.is_available()


def synchronize() -> None:
    """backend.py:124-128"""
    import torch
    torch.cuda.synchronize()


_default_device = None


def select_device(device):
    """backend.py:131-144: explicit device, else the context default, else cpu."""
    return device or _default_device or cpu


@contextmanager
def use_device(device: Device):
    """backend.py:147-166"""
    global _default_device
    prev, _default_device = _default_device, device
    try:
        yield
    finally:
        _default_device = prev


# ---- compute mode of the contractions (Conv2D / Linear)
#   "fp32"      the reference's contract, allclose(rtol=atol=1e-5): fp32-exact on tcgen05 — operands split into tf32 hi + lo
#               planes, three MMAs per k-step, chunked fp32 accumulation (CPT_MODE_FP32X3); layers outside the TMA limits and
#               convolutions with < 8 input channels run the FFMA kernels.  ~11x the FFMA path on the Conv2D sweep.
#   "fp32_simt" every contraction on the exact FFMA kernels (CPT_MODE_FP32): the mode round 1 called "fp32"
#   "tf32"      tcgen05 kind::tf32, 2e-3;   "bf16"  tcgen05 kind::f16 (bf16 operands, fp32 accumulate), 1e-2
#   "fp32x3"    alias of "fp32"
_MODES = {"fp32": _lib.MODE_FP32X3, "fp32x3": _lib.MODE_FP32X3, "fp32_simt": _lib.MODE_FP32, "tf32": _lib.MODE_TF32, "bf16": _lib.MODE_BF16}
_mode = _MODES[os.environ.get("COMPYUTE_B200_MODE", "fp32")]


def set_strip_conv_enabled(enabled: bool) -> bool:
    """Opt-in switch of the strip ("shared halo") convolution kernels for stride-1 same-padded 64 / 128-channel layers in
    bf16 mode (csrc/strip_kernel.cuh).  Returns the previous setting."""
    return bool(_lib.lib().cpt_conv2d_set_strip_enabled(1 if enabled else 0))


def get_compute_mode() -> int:
    return _mode


def set_compute_mode(mode: str | int) -> None:
    global _mode
    _mode = _MODES[mode] if isinstance(mode, str) else int(mode)


@contextmanager
def compute_mode(mode: str | int):
    """``with compute_mode("bf16"):`` — same kind of global switch as ``use_dtype`` (typing.py:133-139)."""
    prev = _mode
    set_compute_mode(mode)
    try:
        yield
    finally:
        set_compute_mode(prev)
