"""Tiny driver for ncu captures: Conv2D C=512 (or $NCU_C), 3x3 same, 56x56, B=256 fwd+bwd in $NCU_MODE (bf16), 3 iterations.
tc_kernel launches per iteration: fprop, dgrad, wgrad (in that order)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200.nn.functional import Conv2DFn, FunctionCache
from compyute_b200.tensors import DeviceArray, Tensor

C, B, mode = int(os.environ.get("NCU_C", 512)), int(os.environ.get("NCU_B", 256)), os.environ.get("NCU_MODE", "bf16")
wrap = lambda t: Tensor(DeviceArray(t, tuple(t.shape), np.float32))
x = wrap(torch.randn(B, C, 56, 56, device="cuda"))
dy = wrap(torch.empty(B, C, 56, 56, device="cuda").uniform_(-0.1, 0.1))
w = wrap(torch.empty(C, C, 3, 3, device="cuda").uniform_(-0.02, 0.02))
b = wrap(torch.zeros(C, device="cuda"))
with cp.compute_mode(mode):
    for _ in range(int(os.environ.get("NCU_ITERS", 3))):
        c = FunctionCache()
        Conv2DFn.forward(c, x, w, b, 1, 1, 1)
        Conv2DFn.backward(c, dy)
torch.cuda.synchronize()
print("done")
