"""CPU tests of the ``compyute`` import shim (repo root ``compyute/``): a program written against dakofler/Compyute finds every
in-scope name under the reference's own module paths; out-of-scope names import but raise when used (out of scope)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve():
    import compyute as cp
    from compyute import nn
    from compyute.backend import Device, cpu, cuda  # noqa: F401
    from compyute.nn import (AvgPooling2D, BatchNorm1D, BatchNorm2D, Conv2D, CrossEntropyLoss, Dropout, Flatten, Linear, MaxPooling2D,  # noqa: F401
                             ReLU, ResidualConnection, Sequential)
    from compyute.nn.modules.convolutions import Conv1D, Conv2D as C2, ConvTranspose1D, ConvTranspose2D  # noqa: F401
    from compyute.nn.optimizers import SGD, Adam, AdamW, NAdam  # noqa: F401
    from compyute.nn.parameter import Buffer, Parameter  # noqa: F401
    from compyute.random.random import seed, set_seed, uniform, uniform_int  # noqa: F401
    from compyute.tensors import ShapeLike, Tensor  # noqa: F401
    from compyute.typing import float32, int64
    import compyute_b200
    assert C2 is compyute_b200.nn.Conv2D and nn.Linear is compyute_b200.nn.Linear and cp.tensor is compyute_b200.tensor
    assert float32 is np.float32 and int64 is np.int64 and cp.bool_ is np.bool_
    for mod in ("activations", "containers", "convolutions", "linear", "module", "normalizations", "poolings", "regularizations", "reshapes"):
        importlib.import_module(f"compyute.nn.modules.{mod}")
    t = cp.tensor(np.arange(6, dtype=np.float32).reshape(2, 3))
    assert t.shape == (2, 3) and t.device == compyute_b200.cpu


def test_out_of_scope_names_raise_when_used():
    from compyute.nn import GELU, LayerNorm, Upsample2D
    from compyute.nn.modules.convolutions import Conv1D
    for cls, args in ((Conv1D, (1, 2, 3)), (LayerNorm, ((4,),)), (GELU, ()), (Upsample2D, (2,))):
        with pytest.raises(NotImplementedError):
            cls(*args)


def test_seed_is_a_context_manager_and_a_decorator():
    from compyute.random.random import seed, uniform

    @seed(42)
    def draw():
        return uniform((4,)).to_numpy()

    a, b = draw(), draw()
    with seed(42):
        c = uniform((4,)).to_numpy()
    assert np.array_equal(a, b) and np.array_equal(a, c)
    np.random.seed(42)
    assert np.array_equal(a, np.random.uniform(0.0, 1.0, (4,)).astype(np.float32))  # host path = the reference's NumPy stream


def test_shim_device_switch_binds_cpu_to_cuda():
    """COMPYUTE_SHIM_DEVICE=cuda (how tests/test_gpu_dropin.py runs the reference's tests): the name `cpu` is the cuda device
    and new tensors default to it — checked in a subprocess, import only (no GPU work)."""
    code = ("import compyute as cp; from compyute.backend import cpu, cuda; import compyute_b200.backend as b; "
            "assert cpu is cuda and b.select_device(None) is cuda; print('ok')")
    env = dict(os.environ, COMPYUTE_SHIM_DEVICE="cuda", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1000:]
