"""Achieved HBM GB/s of every memory-bound kernel on the path, at config-relevant sizes (CUDA events, inputs > L2).
Algorithmic bytes per element follow SURVEY §8d / DESIGN.md §4.  Prints one JSON line per kernel."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200 import _lib, nn
from compyute_b200.nn.functional import (AvgPooling2DFn, BatchNorm2DFn, CrossEntropyLossFn, FunctionCache, MaxPooling2DFn, ReLUFn)
from compyute_b200.tensors import DeviceArray, Tensor

L = _lib.lib()
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
wrap = lambda t: Tensor(DeviceArray(t, tuple(t.shape), np.float32))


def timeit(fn, iters=10, warm=3):
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


def report(name, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "algorithmic_MB": round(nbytes / 1e6, 1), "ms": round(ms, 4), "GB/s": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / PEAK, 3), "note": note}), flush=True)


if __name__ == "__main__":
    B, C, H = 256, 64, 112  # ResNet-18 stem activation: 205.5 M elements = 822 MB
    N = B * C * H * H
    x = wrap(torch.randn(B, C, H, H, device="cuda"))
    dy = wrap(torch.randn(B, C, H, H, device="cuda"))
    c = FunctionCache()
    # ReLU
    y = ReLUFn.forward(c, x)
    report("relu_fwd", 8.125 * N, timeit(lambda: (c.cache.clear(), ReLUFn.forward(c, x))), "8 B/elem + 1/8 B mask")
    c.cache.clear(); ReLUFn.forward(c, x); mask = c.cache[-1]
    def relu_b():
        c.cache.append(mask); ReLUFn.backward(c, dy)
    report("relu_bwd", 8.125 * N, timeit(relu_b))
    # MaxPool k=2
    c.cache.clear()
    report("maxpool2d_fwd k=2", 4 * N + N, timeit(lambda: (c.cache.clear(), MaxPooling2DFn.forward(c, x, 2))), "4N + 4N/k^2")
    c.cache.clear(); yp = MaxPooling2DFn.forward(c, x, 2); saved = c.cache[-1]
    dyp = wrap(torch.randn(B, C, H // 2, H // 2, device="cuda"))
    def mp_b():
        c.cache.append(saved); MaxPooling2DFn.backward(c, dyp)
    report("maxpool2d_bwd k=2", 8 * N + 2 * N, timeit(mp_b), "x read + dx write + y, dy")
    # AvgPool k=7 on (256,512,7,7) is tiny; use k=2 on the big tensor
    report("avgpool2d_fwd k=2", 4 * N + N, timeit(lambda: (c.cache.clear(), AvgPooling2DFn.forward(c, x, 2))))
    # BatchNorm2D
    w = wrap(torch.ones(C, device="cuda")); b = wrap(torch.zeros(C, device="cuda"))
    rm = wrap(torch.zeros(C, device="cuda")); rv = wrap(torch.ones(C, device="cuda"))
    report("bn2d_fwd_train", 12 * N, timeit(lambda: (c.cache.clear(), BatchNorm2DFn.forward(c, x, rm, rv, w, b, 0.1, 1e-5, True))), "x read twice + y write")
    report("bn2d_fwd_eval", 8 * N, timeit(lambda: (c.cache.clear(), BatchNorm2DFn.forward(c, x, rm, rv, w, b, 0.1, 1e-5, False))))
    c.cache.clear(); BatchNorm2DFn.forward(c, x, rm, rv, w, b, 0.1, 1e-5, True); saved = c.cache[-1]
    def bn_b():
        c.cache.append(saved); BatchNorm2DFn.backward(c, dy)
    report("bn2d_bwd", 20 * N, timeit(bn_b), "dy, x read twice + dx write")
    # residual add / grad accumulation
    a = DeviceArray(torch.randn(N, device="cuda"), (N,), np.float32); b2 = DeviceArray(torch.randn(N, device="cuda"), (N,), np.float32)
    def add():
        nonlocal_a = a
        nonlocal_a += b2
    report("add_inplace", 12 * N, timeit(add))
    report("isnan_flag", 4 * N, timeit(lambda: L.cpt_isnan_flag(a.ptr, N, DeviceArray.zeros((1,), np.int32).ptr, cp.tensors.stream_ptr())))
    # NCHW -> NHWC staging (bf16 / tf32), with and without the fused channel sum
    for mode, bpe in ((_lib.MODE_BF16, 6), (_lib.MODE_TF32, 8)):
        dst = DeviceArray.empty((L.cpt_channels_last_bytes(B, C, H, H, mode),), np.uint8)
        cs = DeviceArray.zeros((C,), np.float32)
        wsz = L.cpt_to_channels_last_workspace_size(B, C, H, H); wsb = DeviceArray.empty((wsz,), np.uint8)
        report(f"nchw_to_nhwc mode={mode}", bpe * N, timeit(lambda: L.cpt_to_channels_last(x.data.ptr, dst.ptr, B, C, H, H, mode, None, None, 0, cp.tensors.stream_ptr())))
        report(f"nchw_to_nhwc+chan_sum mode={mode}", bpe * N, timeit(lambda: L.cpt_to_channels_last(x.data.ptr, dst.ptr, B, C, H, H, mode, cs.ptr, wsb.ptr, wsz, cp.tensors.stream_ptr())))
    del x, dy, y, a, b2
    torch.cuda.empty_cache()
    # Adam / SGD on the MLP config's parameter count (8 x 4096^2 + biases = 134 M)
    ps = [nn.Parameter(wrap(torch.randn(4096, 4096, device="cuda"))) for _ in range(8)]
    for p in ps:
        p.grad = wrap(torch.randn(4096, 4096, device="cuda"))
    P = sum(p.size for p in ps)
    adam = nn.optimizers.Adam(ps, lr=1e-3)
    adam.step()
    report("adam_step (134M params, 1 launch)", 28 * P, timeit(adam.step), "r p,g,m,v  w p,m,v")
    sgd = nn.optimizers.SGD(ps, lr=1e-3, momentum=0.9)
    sgd.step()
    report("sgd_momentum_step", 20 * P, timeit(sgd.step))
    # softmax-CE on the MLP logits (8192 x 4096)
    logits = wrap(torch.randn(8192, 4096, device="cuda")); t = Tensor(DeviceArray(torch.randint(0, 4096, (8192,), dtype=torch.int32, device="cuda"), (8192,), np.int32))
    cc = FunctionCache()
    n = 8192 * 4096
    report("softmax_ce_fwd 8192x4096", 8 * n, timeit(lambda: (cc.cache.clear(), CrossEntropyLossFn.forward(cc, logits, t, 1e-8))), "logits read + probs write (sub-L2 size)")
    del logits, ps, adam, sgd
    torch.cuda.empty_cache()
    # generic device-tensor operators (csrc/tensor_ops.cu) on the ResNet-18 stem activation shape
    xs = wrap(torch.randn(B, C, H, H, device="cuda")); ys = wrap(torch.randn(B, C, H, H, device="cuda"))
    ch = wrap(torch.randn(C, 1, 1, device="cuda"))
    report("ew_binary add (same shape)", 12 * N, timeit(lambda: xs + ys), "a, b read + out write")
    report("ew_binary mul scalar", 8 * N, timeit(lambda: xs * 1.5))
    report("ew_binary add (C,1,1) broadcast", 8 * N, timeit(lambda: xs + ch), "strided index path")
    report("ew_binary gt -> bool", 5 * N, timeit(lambda: xs > 0.5), "4 B read + 1 B write")
    report("ew_unary exp", 8 * N, timeit(lambda: cp.exp(xs)))
    report("reduce sum (all)", 4 * N, timeit(lambda: xs.sum()), "row form, split over the grid + fixed-order finish")
    report("reduce sum dims (0,2,3)", 4 * N, timeit(lambda: xs.sum((0, 2, 3))), "row form, 2 reduced dims")
    report("reduce max last dim", 4 * N, timeit(lambda: xs.max(-1)), "row form, 112-element rows")
    x2 = wrap(torch.randn(8192 * 8, 4096, device="cuda")); n2 = x2.size
    report("reduce sum dim 0 of (65536,4096)", 4 * n2, timeit(lambda: x2.sum(0)), "column form")
    report("reduce argmax dim 1 of (65536,4096)", 4 * n2, timeit(lambda: x2.argmax(1)))
    report("permute NCHW->NHWC (strided copy)", 8 * N, timeit(lambda: xs.permute((0, 2, 3, 1))), "generic gather; the staging kernel is the tuned form")
    report("slice copy x[:, :, 1:-1, 1:-1]", 8 * (B * C * (H - 2) ** 2), timeit(lambda: xs[:, :, 1:-1, 1:-1]))
    idx = Tensor(DeviceArray(torch.randperm(B, device="cuda").to(torch.int32), (B,), np.int32))
    report("gather_rows (batch shuffle)", 8 * N, timeit(lambda: xs[idx]))
    report("cast f32 -> int32", 8 * N, timeit(lambda: xs.to_int()))
    report("random normal", 4 * N, timeit(lambda: cp.random.normal((B, C, H, H), device=cp.cuda)), "write only")
