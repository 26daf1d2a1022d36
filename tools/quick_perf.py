"""Quick device-time probe of the Conv2D passes at BASELINE config-2 shapes (CUDA events, warm-up, L2-sized inputs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200 import _lib
from compyute_b200.nn.functional import Conv2DFn, FunctionCache

T = lambda a: cp.tensor(a, device=cp.cuda)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    B, H, K = int(os.environ.get("QP_B", 256)), 56, 3
    rng = np.random.RandomState(0)
    for C in (64, 128, 256, 512):
        x = T(rng.uniform(-0.1, 0.1, (B, C, H, H)).astype(np.float32))
        w = T(rng.uniform(-0.05, 0.05, (C, C, K, K)).astype(np.float32))
        b = T(rng.uniform(-0.05, 0.05, (C,)).astype(np.float32))
        dy = T(rng.uniform(-0.1, 0.1, (B, C, H, H)).astype(np.float32))
        flops = 2.0 * B * C * C * H * H * K * K
        for mode in ("bf16", "tf32", "fp32"):
            if mode == "fp32" and C > 128 and B > 64:
                continue
            with cp.compute_mode(mode):
                c = FunctionCache()

                def fwd():
                    c.cache.clear()
                    return Conv2DFn.forward(c, x, w, b, 1, 1, 1)

                def fwdbwd():
                    c.cache.clear()
                    Conv2DFn.forward(c, x, w, b, 1, 1, 1)
                    Conv2DFn.backward(c, dy)

                tf = timeit(fwd)
                tfb = timeit(fwdbwd)
            st = _lib.lib().cpt_tc_check_status()
            print(f"C={C:4d} {mode:5s} fwd {tf:8.3f} ms ({flops / tf / 1e9:8.1f} TFLOP/s)  fwd+bwd {tfb:8.3f} ms "
                  f"({3 * flops / tfb / 1e9:8.1f} TFLOP/s) status={st}", flush=True)
