"""Device time of the ResNet stem convolution (3 -> 64, 7x7 / s2 / p3, 224x224, B = 256, bf16 packed-K path): forward
(im2col_pack + GEMM) and backward (staging of dy, GEMM -> dcol, col2im, wgrad), CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200.nn.functional import Conv2DFn, FunctionCache
from tools.quick_perf import timeit

if __name__ == "__main__":
    B = int(os.environ.get("QP_B", 256))
    rng = np.random.RandomState(0)
    T = lambda a: cp.tensor(a, device=cp.cuda)
    x = T(rng.normal(0, 1, (B, 3, 224, 224)).astype(np.float32))
    w = T(rng.uniform(-0.05, 0.05, (64, 3, 7, 7)).astype(np.float32))
    dy = T(rng.uniform(-0.1, 0.1, (B, 64, 112, 112)).astype(np.float32))
    with cp.compute_mode("bf16"):
        c = FunctionCache()

        def fwd():
            c.cache.clear()
            return Conv2DFn.forward(c, x, w, None, 3, 2, 1)

        def fwdbwd():
            c.cache.clear()
            Conv2DFn.forward(c, x, w, None, 3, 2, 1)
            return Conv2DFn.backward(c, dy)

        tf, tfb = timeit(fwd, 10, 3), timeit(fwdbwd, 10, 3)
        dx = fwdbwd()[0].to_numpy()
    print(f"stem fwd {tf:.3f} ms, bwd {tfb - tf:.3f} ms, dx checksum {float(np.abs(dx).sum()):.6e}")
