"""Diagnostic for the tcgen05 path (run on the GPU box): exact integer-valued problems per operand layout, so a
descriptor / swizzle / im2col mistake shows up as a precise mismatch pattern.  Continues after failures."""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200 import _lib
from compyute_b200.nn.functional import Conv2DFn, FunctionCache, LinearFn
from oracle import compyute_ref as R

L = _lib.lib()
T = lambda a: cp.tensor(a, device=cp.cuda)
rng = np.random.RandomState(0)


def ints(shape, lo=-2, hi=3):
    return rng.randint(lo, hi, shape).astype(np.float32)


def report(name, got, ref):
    got = got.to_numpy()
    bad = got != ref
    st = L.cpt_tc_check_status()
    msg = f"{name:44s} status={st} mismatches={bad.sum()}/{bad.size} maxerr={np.abs(got - ref).max():.4g}"
    if bad.any():
        idx = np.argwhere(bad)
        msg += f" first_bad={idx[0].tolist()} last_bad={idx[-1].tolist()} got={got[tuple(idx[0])]} ref={ref[tuple(idx[0])]}"
        # which rows/cols are affected
        for ax in range(got.ndim):
            others = tuple(a for a in range(got.ndim) if a != ax)
            frac = bad.mean(axis=others)
            nz = np.nonzero(frac)[0]
            msg += f"\n      axis{ax}: bad idx count {len(nz)}/{got.shape[ax]} e.g. {nz[:12].tolist()}"
    print(msg, flush=True)
    return not bad.any() and st == 0


def run(name, fn):
    try:
        return fn()
    except Exception:
        print(f"{name}: EXCEPTION\n{traceback.format_exc()}", flush=True)
        try:
            L.cpt_tc_check_status()
        except Exception:
            pass
        return False


def linear_case(mode, N, In, Out):
    def f():
        x, w, b, dy = ints((N, In)), ints((Out, In)), ints((Out,)), ints((N, Out))
        rc = []
        y_ref = R.linear_forward(rc, x, w, b)
        dx_ref, dw_ref, db_ref = R.linear_backward(rc, dy)
        ok = True
        with cp.compute_mode(mode):
            c = FunctionCache()
            y = LinearFn.forward(c, T(x), T(w), T(b))
            ok &= report(f"linear fwd   {mode} N={N} In={In} Out={Out}", y, y_ref)
            dx, dw, db = LinearFn.backward(c, T(dy))
            ok &= report(f"linear dgrad {mode} N={N} In={In} Out={Out}", dx, dx_ref)
            ok &= report(f"linear wgrad {mode} N={N} In={In} Out={Out}", dw, dw_ref)
            ok &= report(f"linear db    {mode} N={N} In={In} Out={Out}", db, db_ref)
        return ok
    return run(f"linear {mode} {N} {In} {Out}", f)


def conv_case(mode, B, Ci, Co, H, K, P, s, d):
    def f():
        x, w, b = ints((B, Ci, H, H)), ints((Co, Ci, K, K), -1, 2), ints((Co,))
        rc = []
        y_ref = R.conv2d_forward(rc, x, w, b, P, s, d)
        dy = ints(y_ref.shape, -1, 2)
        dx_ref, dw_ref, db_ref = R.conv2d_backward(rc, dy)
        ok = True
        tag = f"{mode} B{B} Ci{Ci} Co{Co} H{H} K{K} P{P} s{s} d{d}"
        with cp.compute_mode(mode):
            c = FunctionCache()
            y = Conv2DFn.forward(c, T(x), T(w), T(b), P, s, d)
            ok &= report(f"conv fprop {tag}", y, y_ref)
            dx, dw, db = Conv2DFn.backward(c, T(dy))
            ok &= report(f"conv dgrad {tag}", dx, dx_ref)
            ok &= report(f"conv wgrad {tag}", dw, dw_ref)
            ok &= report(f"conv db    {tag}", db, db_ref)
        return ok
    return run(f"conv {mode}", f)


if __name__ == "__main__":
    import ctypes
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.cpt_device_info(0, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), None, None)
    print(f"device: {sm.value} SMs, cc {ma.value}.{mi.value}", flush=True)
    results = []
    for mode in ("fp32", "tf32", "bf16"):
        results.append(linear_case(mode, 128, 64, 128))      # single tile, two k-iterations (tf32) / one (bf16)
        results.append(linear_case(mode, 256, 256, 256))     # multi tile
        results.append(linear_case(mode, 300, 200, 136))     # ragged edges (TMA zero fill)
        results.append(linear_case(mode, 1024, 512, 384))
        results.append(conv_case(mode, 2, 64, 64, 8, 3, 1, 1, 1))    # one M tile (128 pixels), im2col halo
        results.append(conv_case(mode, 2, 64, 64, 8, 1, 0, 1, 1))    # 1x1: im2col without taps
        results.append(conv_case(mode, 3, 32, 48, 12, 3, 1, 1, 1))   # ragged channels / pixels
        results.append(conv_case(mode, 2, 128, 256, 16, 3, 1, 1, 1))
        results.append(conv_case(mode, 2, 16, 32, 12, 5, 2, 1, 1))
        results.append(conv_case(mode, 2, 32, 32, 16, 3, 1, 2, 1))   # strided fprop / wgrad (dgrad -> exact path)
        results.append(conv_case(mode, 2, 16, 16, 12, 3, 2, 1, 2))   # dilation
        results.append(conv_case(mode, 2, 3, 8, 10, 3, 1, 1, 1))     # tiny channel count
        results.append(conv_case(mode, 2, 64, 128, 16, 1, 0, 2, 1))  # 1x1 stride 2: empty stride classes (dx = 0 there)
        results.append(conv_case(mode, 2, 3, 16, 30, 7, 3, 2, 1))    # ResNet stem shape: 7x7 / s2 / p3, Ci = 3
        results.append(conv_case(mode, 2, 16, 24, 17, 3, 1, 3, 1))   # stride 3, odd extent
        results.append(conv_case(mode, 1, 8, 8, 13, 3, 0, 2, 2))     # stride 2 + dilation 2
    print("SUMMARY", sum(bool(r) for r in results), "/", len(results), flush=True)
