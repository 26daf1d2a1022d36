"""Probe: how does the tcgen05 fp32 accumulator round?  Inputs exactly representable in tf32 (11-bit significands) and all
positive, so every product is exact in fp32 and the only error source is the accumulation.  Round-to-nearest gives a
zero-mean error ~ sqrt(N) ulp; truncation gives a negative bias growing ~ N ulp."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200.nn.functional import linear

rng = np.random.RandomState(0)
for K in (256, 1024, 4096, 16384, 65536):
    N, Out = 256, 128
    x = (1.0 + rng.randint(0, 1024, (N, K)) / 1024.0).astype(np.float32)      # in [1, 2), 10 fractional bits
    w = (1.0 + rng.randint(0, 1024, (Out, K)) / 1024.0).astype(np.float32)
    exact = x.astype(np.float64) @ w.astype(np.float64).T
    with cp.compute_mode("tf32"):
        y = linear(cp.tensor(x, device=cp.cuda), cp.tensor(w, device=cp.cuda)).to_numpy().astype(np.float64)
    f32 = (x @ w.T).astype(np.float64)  # numpy/OpenBLAS fp32 for comparison
    rel = (y - exact) / exact
    relf = (f32 - exact) / exact
    print(f"K={K:6d}  tcgen05 tf32: mean rel err {rel.mean():+.3e}  max |rel| {np.abs(rel).max():.3e}   |  numpy fp32: mean {relf.mean():+.3e} max {np.abs(relf).max():.3e}", flush=True)
