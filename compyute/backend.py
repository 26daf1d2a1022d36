"""compyute/backend.py of the reference: devices (``cpu``, ``cuda``), ``use_device``, availability helpers."""

import os as _os

from compyute_b200 import backend as _b
from compyute_b200.backend import *  # noqa: F401,F403
from compyute_b200.backend import Device, DeviceError, cuda, gpu_available, select_device, synchronize, use_device  # noqa: F401

CUDARuntimeError = __import__("compyute_b200._lib", fromlist=["CudaRuntimeError"]).CudaRuntimeError

if _os.environ.get("COMPYUTE_SHIM_DEVICE", "") == "cuda":
    cpu = cuda  # the reference's tests pass device=cpu explicitly: route them to the only device this package computes on
    _b._default_device = cuda
else:
    cpu = _b.cpu


def free_cuda_memory() -> None:
    """backend.py:117-121 of the reference (returns cached blocks to the driver)."""
    import torch
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
