"""``compyute`` — import shim over :mod:`compyute_b200` for the hot path (BASELINE north_star: "drop-in replacement").

A program written against dakofler/Compyute (``import compyute as cp``, ``from compyute import nn``,
``from compyute.nn.modules.convolutions import Conv2D`` ...) imports THIS package when the repository root precedes the
reference on ``sys.path`` and gets the B200 implementation of every in-scope name under the reference's own module paths
(compyute/__init__.py:6-14 of the reference).  Names outside SURVEY §8 (Conv1D, LayerNorm, GELU, ...) resolve to stubs that
raise ``NotImplementedError`` when instantiated, so that reference modules which import them by name still import.

``COMPYUTE_SHIM_DEVICE=cuda`` (what tests/test_gpu_dropin.py sets to run the reference's own unit tests on the GPU): the
default device becomes ``cuda`` and the name ``cpu`` is bound to the cuda device, because the reference's test utilities
hard-code ``device=cpu`` (tests/utils.py:21) — there is no CPU path to fall back to here.
"""

import os as _os

import compyute_b200 as _impl
from compyute_b200 import *  # noqa: F401,F403
from compyute_b200 import DeviceArray, ShapeError, Tensor, tensor  # noqa: F401

from . import backend, nn, random, tensors, typing  # noqa: F401,E402
from .backend import *  # noqa: F401,F403,E402
from .typing import *  # noqa: F401,F403,E402

__version__ = "0.1.8+b200." + _impl.__version__
