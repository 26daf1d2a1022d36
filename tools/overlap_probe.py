"""Does the NCHW->NHWC staging kernel (HBM-bound) co-run with the persistent tcgen05 GEMM (tensor-bound)?
Times, with CUDA events on the main stream: (a) staging alone, (b) fprop alone, (c) both serial on one stream,
(d) fprop on the main stream with staging of ANOTHER buffer on a side stream.  If (d) ~ max(a, b) the two kernels share
the SMs (the GEMM leaves ~34 KB of shared memory per SM) and a chunked stage/GEMM pipeline hides the staging pass."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp  # noqa: E402,F401
from compyute_b200 import _lib  # noqa: E402

L = _lib.lib()
C, B, H = int(os.environ.get("PROBE_C", 512)), int(os.environ.get("PROBE_B", 256)), 56
mode = _lib.MODE_BF16
d = _lib.ConvDesc(B, C, H, H, C, 3, 1, 1, 1)
x = torch.randn(B, C, H, H, device="cuda")
x2 = torch.randn(B, C, H, H, device="cuda")
w = torch.empty(C, C, 3, 3, device="cuda").uniform_(-0.02, 0.02)
bias = torch.zeros(C, device="cuda")
y = torch.empty(B, C, H, H, device="cuda")
nb = L.cpt_channels_last_bytes(B, C, H, H, mode)
x_cl = torch.empty(nb, dtype=torch.uint8, device="cuda")
x2_cl = torch.empty(nb, dtype=torch.uint8, device="cuda")
wsb = L.cpt_conv2d_workspace_size(_lib.OP_FPROP, ctypes.byref(d), mode)
ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device="cuda")
main = torch.cuda.current_stream()
side = torch.cuda.Stream()
P = lambda t: ctypes.c_void_p(t.data_ptr())


def stage(src, dst, st):
    _lib.check(L.cpt_to_channels_last(P(src), P(dst), B, C, H, H, mode, None, None, 0, st.cuda_stream))


def fprop(st):
    _lib.check(L.cpt_conv2d_fprop_cl(ctypes.byref(d), P(x_cl), P(w), P(bias), P(y), mode, P(ws), wsb, st.cuda_stream))


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(n):
        fn()
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def both_serial():
    stage(x2, x2_cl, main)
    fprop(main)


def both_overlap():
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    fprop(main)
    stage(x2, x2_cl, side)
    ev2 = torch.cuda.Event()
    ev2.record(side)
    main.wait_event(ev2)


def both_overlap_stage_first():
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    stage(x2, x2_cl, side)
    fprop(main)
    ev2 = torch.cuda.Event()
    ev2.record(side)
    main.wait_event(ev2)


stage(x, x_cl, main)
torch.cuda.synchronize()
res = {"C": C, "B": B,
       "stage_ms": timed(lambda: stage(x2, x2_cl, main)),
       "fprop_ms": timed(lambda: fprop(main)),
       "serial_ms": timed(both_serial),
       "overlap_gemm_first_ms": timed(both_overlap),
       "overlap_stage_first_ms": timed(both_overlap_stage_first)}
print(res)
assert L.cpt_tc_check_status() == 0
