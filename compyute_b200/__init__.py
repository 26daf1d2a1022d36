"""compyute_b200 — B200-native CNN-training hot path behind Compyute's nn.functional / nn.Module API.

    import compyute_b200 as cp
    from compyute_b200 import nn
    with cp.use_device(cp.cuda):
        model = nn.Sequential(nn.Conv2D(3, 64, 3, padding="same"), nn.ReLU(), ...)

Hand-written CUDA for sm_100a (compyute_b200/csrc) behind a C ABI (include/compyute_b200.h); no CPU fallback.
"""

from . import distributed, graph, nn, random, tensor_ops
from .backend import *
from .tensors import DeviceArray, ShapeError, Tensor, tensor
from .utils import load, save
from .tensor_ops import *  # noqa: F401,F403  (cp.zeros, cp.exp, cp.sum, cp.concat, ... like the reference's top level)

__version__ = "0.1.0"
