"""Which operand rounding do the tensor-core modes apply?  Device results of Linear / Conv2D (fwd, dgrad, wgrad) in tf32 and bf16
mode against fp64 contractions of operands rounded on the host in several ways; the matching emulation agrees to ~1e-6 of the
largest magnitude, the others to ~1e-4 (tf32) / 1e-3 (bf16).  Pins oracle/model_ref.round_operand (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200.nn.functional import Conv2DFn, FunctionCache, LinearFn
from oracle.model_ref import round_operand


def rne_tf32(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


VARIANTS = {"none": lambda a: a, "tf32_trunc": lambda a: round_operand(a, "tf32_trunc"), "tf32_rna": lambda a: round_operand(a, "tf32_rna"),
            "tf32_rne": rne_tf32, "bf16": lambda a: round_operand(a, "bf16")}
T = lambda a: cp.tensor(a, device=cp.cuda)
rel = lambda got, ref: float(np.abs(got - ref).max() / np.abs(ref).max())
rng = np.random.RandomState(0)

for mode in ("tf32", "bf16"):
    N, In, Out = 256, 512, 256
    x, w, dy = (rng.normal(0, 1, s).astype(np.float32) for s in ((N, In), (Out, In), (N, Out)))
    with cp.compute_mode(mode):
        c = FunctionCache()
        y = LinearFn.forward(c, T(x), T(w), None).to_numpy()
        dx, dw, _ = LinearFn.backward(c, T(dy))
    dx, dw = dx.to_numpy(), dw.to_numpy()
    for name, f in VARIANTS.items():
        xr, wr, gr = (f(a).astype(np.float64) for a in (x, w, dy))
        print(f"linear {mode:5s} emulation {name:10s}: y {rel(y, xr @ wr.T):.2e}  dx {rel(dx, gr @ wr):.2e}  dw {rel(dw, gr.T @ xr):.2e}", flush=True)
    B, C, H = 4, 64, 16
    x, w, dy = (rng.normal(0, 1, s).astype(np.float32) for s in ((B, C, H, H), (C, C, 3, 3), (B, C, H, H)))
    with cp.compute_mode(mode):
        c = FunctionCache()
        y = Conv2DFn.forward(c, T(x), T(w), None, 1, 1, 1).to_numpy()
        dx, dw, _ = Conv2DFn.backward(c, T(dy))
    dx, dw = dx.to_numpy(), dw.to_numpy()
    for name, f in VARIANTS.items():
        xt = torch.from_numpy(f(x).astype(np.float64)).requires_grad_(True)
        wt = torch.from_numpy(f(w).astype(np.float64)).requires_grad_(True)
        yt = torch.nn.functional.conv2d(xt, wt, padding=1)
        yt.backward(torch.from_numpy(f(dy).astype(np.float64)))
        print(f"conv   {mode:5s} emulation {name:10s}: y {rel(y, yt.detach().numpy()):.2e}  dx {rel(dx, xt.grad.numpy()):.2e}  dw {rel(dw, wt.grad.numpy()):.2e}", flush=True)
