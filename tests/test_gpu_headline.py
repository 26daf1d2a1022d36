"""Parity of the kernels the headline benchmark actually times (``-m gpu``): BASELINE configs[1] at C in {128, 256, 512},
56x56, batch 256 — i.e. ``tc_kernel<BF16|TF32, BN=256, OP_CONV, CTA2>`` (fprop / dgrad) and ``tc_kernel<.., MN, MN, BN=256,
OP_WGRAD, CTA2>`` (wgrad), plus every operand layout of the tensor-core kernel on exact-integer problems.

Integer-valued operands (|x| <= 2, |w| <= 1, |dy| <= 1) are exactly representable in bf16 / tf32 and every partial sum stays
below 2^24, so ANY correct evaluation order gives the same fp32 bits: all comparisons here are ``==``.  References:
  * small shapes — the CPU oracle (oracle/compyute_ref.py: the reference's as_strided + einsum algorithm,
    convolution_funcs.py:357-410);
  * full size — an fp64 CPU convolution of a subset of the images (y, dx: complete planes), fp64 dot products of sampled
    filter-gradient entries over the whole batch (dw), dy.sum (db), and the repo's exact FFMA path on the whole tensors.
"""
import numpy as np
import pytest

from oracle import compyute_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cp():
    import compyute_b200 as cp
    from compyute_b200 import _lib
    _lib.lib()
    assert cp.gpu_available(), "no CUDA device"
    return cp


def tc_ok():
    from compyute_b200 import _lib
    assert _lib.lib().cpt_tc_check_status() == 0, "tensor-core pipeline watchdog fired"


def ints(rng, shape, lo, hi):
    return rng.randint(lo, hi + 1, shape, dtype=np.int8).astype(np.float32)


def exact(name, got, ref):
    got = got.to_numpy() if hasattr(got, "to_numpy") else got
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    bad = got != ref
    assert not bad.any(), f"{name}: {int(bad.sum())}/{bad.size} mismatches, first at {np.argwhere(bad)[0].tolist()}, max err {np.abs(got - ref).max():.4g}"


# ------------------------------------------------------------------ every operand layout, exact (was tools/tc_diag.py)
LINEAR_EXACT = [(128, 64, 128), (256, 256, 256), (300, 200, 136), (1024, 512, 384), (512, 1024, 512)]
CONV_EXACT = [  # (B, Ci, Co, H, K, pad, stride, dil)
    (2, 64, 64, 8, 3, 1, 1, 1),      # one M tile (128 pixels), im2col halo
    (2, 64, 64, 8, 1, 0, 1, 1),      # 1x1
    (3, 32, 48, 12, 3, 1, 1, 1),     # ragged channels / pixels
    (2, 128, 256, 16, 3, 1, 1, 1),   # BN=256, 2-CTA fprop (512 pixels = 2 tile pairs); wgrad BN=256 1-CTA
    (2, 256, 256, 16, 3, 1, 1, 1),   # wgrad <MN, MN, BN=256, OP_WGRAD, CTA2> (Ci = 256: two 128-lane tiles)
    (1, 512, 512, 16, 3, 1, 1, 1),   # the C=512 kernels of the sweep: two N tiles, 72 k-iterations per tile
    (2, 16, 32, 12, 5, 2, 1, 1),
    (2, 32, 32, 16, 3, 1, 2, 1),     # strided fprop / wgrad
    (2, 16, 16, 12, 3, 2, 1, 2),     # dilation
    (2, 3, 8, 10, 3, 1, 1, 1),       # tiny channel count (packed-K path in bf16 mode)
    (2, 64, 128, 16, 1, 0, 2, 1),    # 1x1 stride 2: empty stride classes
    (2, 3, 16, 30, 7, 3, 2, 1),      # ResNet stem shape
    (2, 16, 24, 17, 3, 1, 3, 1),     # stride 3, odd extent
    (1, 8, 8, 13, 3, 0, 2, 2),       # stride 2 + dilation 2
    (4, 64, 64, 56, 3, 1, 1, 1),     # small-C layer of the sweep (strip kernel when enabled)
    (2, 128, 128, 28, 3, 1, 1, 1),
]


@pytest.mark.parametrize("mode", ["fp32", "fp32_simt", "tf32", "bf16"])
@pytest.mark.parametrize("shape", LINEAR_EXACT, ids=lambda s: "x".join(map(str, s)))
def test_linear_exact_integer(cp, shape, mode):
    from compyute_b200.nn.functional import FunctionCache, LinearFn
    N, In, Out = shape
    rng = np.random.RandomState(0)
    x, w, b, dy = ints(rng, (N, In), -2, 2), ints(rng, (Out, In), -2, 2), ints(rng, (Out,), -2, 2), ints(rng, (N, Out), -2, 2)
    rc = []
    y_ref = R.linear_forward(rc, x, w, b)
    dx_ref, dw_ref, db_ref = R.linear_backward(rc, dy)
    T = lambda a: cp.tensor(a, device=cp.cuda)
    with cp.compute_mode(mode):
        c = FunctionCache()
        y = LinearFn.forward(c, T(x), T(w), T(b))
        dx, dw, db = LinearFn.backward(c, T(dy))
    tc_ok()
    exact("y", y, y_ref); exact("dx", dx, dx_ref); exact("dw", dw, dw_ref); exact("db", db, db_ref)


def _torch_conv_ref(x, w, b, dy, P, s, d):
    """fp64 CPU convolution + its gradients (exact on integer-valued data): independent of the repo and of the oracle."""
    import torch
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    wt = torch.from_numpy(w.astype(np.float64)).requires_grad_(True)
    bt = torch.from_numpy(b.astype(np.float64)).requires_grad_(True) if b is not None else None
    y = torch.nn.functional.conv2d(xt, wt, bt, stride=s, padding=P, dilation=d)
    y.backward(torch.from_numpy(dy.astype(np.float64)))
    f = lambda t: t.detach().numpy().astype(np.float32)
    return f(y), f(xt.grad), f(wt.grad), (f(bt.grad) if bt is not None else None)


@pytest.mark.parametrize("mode", ["fp32", "fp32_simt", "tf32", "bf16"])
@pytest.mark.parametrize("shape", CONV_EXACT, ids=lambda s: "x".join(map(str, s)))
def test_conv2d_exact_integer(cp, shape, mode):
    from compyute_b200.nn.functional import Conv2DFn, FunctionCache
    B, Ci, Co, H, K, P, s, d = shape
    rng = np.random.RandomState(0)
    x, w, b = ints(rng, (B, Ci, H, H), -2, 2), ints(rng, (Co, Ci, K, K), -1, 1), ints(rng, (Co,), -2, 2)
    flops = 2.0 * B * Co * Ci * H * H * K * K / (s * s)
    if flops <= 1.5e9:  # the einsum path runs at ~0.5 GFLOP/s: keep the oracle for what it finishes in seconds
        rc = []
        y_ref = R.conv2d_forward(rc, x, w, b, P, s, d)
        dy = ints(rng, y_ref.shape, -1, 1)
        dx_ref, dw_ref, db_ref = R.conv2d_backward(rc, dy)
        t_y, t_dx, t_dw, t_db = _torch_conv_ref(x, w, b, dy, P, s, d)
        # the two references agree bit for bit on integer data — pins the fp64 convolution used for the large cases below
        assert np.array_equal(t_y, y_ref) and np.array_equal(t_dx, dx_ref) and np.array_equal(t_dw, dw_ref) and np.array_equal(t_db, db_ref)
    else:
        Ho = (H + 2 * P - d * (K - 1) - 1) // s + 1
        dy = ints(rng, (B, Co, Ho, Ho), -1, 1)
        y_ref, dx_ref, dw_ref, db_ref = _torch_conv_ref(x, w, b, dy, P, s, d)
    T = lambda a: cp.tensor(a, device=cp.cuda)
    with cp.compute_mode(mode):
        c = FunctionCache()
        y = Conv2DFn.forward(c, T(x), T(w), T(b), P, s, d)
        dx, dw, db = Conv2DFn.backward(c, T(dy))
    tc_ok()
    exact("y", y, y_ref); exact("dx", dx, dx_ref); exact("dw", dw, dw_ref); exact("db", db, db_ref)


@pytest.mark.parametrize("shape", [(4, 64, 64, 56, 3, 1, 1, 1), (2, 128, 128, 28, 3, 1, 1, 1), (3, 64, 128, 20, 3, 1, 1, 1), (2, 128, 64, 17, 5, 2, 1, 1),
                                   (1, 64, 64, 112, 3, 1, 1, 1)], ids=lambda s: "x".join(map(str, s)))
def test_conv2d_strip_kernels_exact_integer(cp, shape):
    """The opt-in strip ("shared halo") kernels (csrc/strip_kernel.cuh): zero-padded channels-last staging, one strip per 128
    outputs with the K*K taps as descriptors into it, resident / streamed filter tiles, wgrad through im2col maps over the
    padded tensors — exact on integer data against the fp64 convolution, like every other operand layout."""
    from compyute_b200.nn.functional import Conv2DFn, FunctionCache
    B, Ci, Co, H, K, P, s, d = shape
    rng = np.random.RandomState(1)
    x, w, b = ints(rng, (B, Ci, H, H), -2, 2), ints(rng, (Co, Ci, K, K), -1, 1), ints(rng, (Co,), -2, 2)
    dy = ints(rng, (B, Co, H, H), -1, 1)
    y_ref, dx_ref, dw_ref, db_ref = _torch_conv_ref(x, w, b, dy, P, s, d)
    T = lambda a: cp.tensor(a, device=cp.cuda)
    prev = cp.set_strip_conv_enabled(True)
    try:
        from compyute_b200 import _lib
        import ctypes
        desc = _lib.ConvDesc(B, Ci, H, H, Co, K, P, s, d)
        assert _lib.lib().cpt_conv2d_strip_supported(ctypes.byref(desc), _lib.MODE_BF16) == 1
        with cp.compute_mode("bf16"):
            c = FunctionCache()
            y = Conv2DFn.forward(c, T(x), T(w), T(b), P, s, d)
            dx, dw, db = Conv2DFn.backward(c, T(dy))
    finally:
        cp.set_strip_conv_enabled(prev)
    tc_ok()
    exact("y", y, y_ref); exact("dx", dx, dx_ref); exact("dw", dw, dw_ref); exact("db", db, db_ref)


# ------------------------------------------------------------------ BASELINE configs[1] at full size
@pytest.mark.parametrize("mode", ["bf16", "tf32", "fp32", "fp32_simt"])
@pytest.mark.parametrize("C", [64, 128, 256, 512])
def test_conv2d_sweep_full_size_exact(cp, C, mode):
    """The benchmarked shapes themselves: x = (256, C, 56, 56), 3x3 same.  The FFMA mode is limited to C <= 128 (it needs
    seconds per pass at C = 512) and runs at full size as the whole-tensor cross-check of the bf16 case."""
    import torch
    from compyute_b200.nn.functional import Conv2DFn, FunctionCache
    if mode == "fp32_simt" and C > 128:
        pytest.skip("the FFMA path at C >= 256 runs as the whole-tensor cross-check of the bf16 case")
    B, H, K = 256, 56, 3
    rng = np.random.RandomState(C)
    x, w, b, dy = ints(rng, (B, C, H, H), -2, 2), ints(rng, (C, C, K, K), -1, 1), ints(rng, (C,), -2, 2), ints(rng, (B, C, H, H), -1, 1)
    T = lambda a: cp.tensor(a, device=cp.cuda)
    xt, wt, bt, dyt = T(x), T(w), T(b), T(dy)
    with cp.compute_mode(mode):
        c = FunctionCache()
        y = Conv2DFn.forward(c, xt, wt, bt, 1, 1, 1)
        dx, dw, db = Conv2DFn.backward(c, dyt)
    tc_ok()
    yh, dxh, dwh, dbh = y.to_numpy(), dx.to_numpy(), dw.to_numpy(), db.to_numpy()
    # (1) complete planes of a few images against an fp64 CPU convolution
    sel = [0, 101, B - 1]
    xs = torch.from_numpy(x[sel].astype(np.float64))
    ws = torch.from_numpy(w.astype(np.float64))
    y_ref = torch.nn.functional.conv2d(xs, ws, torch.from_numpy(b.astype(np.float64)), padding=1).numpy().astype(np.float32)
    exact("y[sel]", yh[sel], y_ref)
    dx_ref = torch.nn.functional.conv_transpose2d(torch.from_numpy(dy[sel].astype(np.float64)), ws, padding=1).numpy().astype(np.float32)
    exact("dx[sel]", dxh[sel], dx_ref)
    # (2) sampled filter-gradient entries: dot products over the whole batch (802816 terms each)
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    for _ in range(24):
        o, i, j, k = rng.randint(C), rng.randint(C), rng.randint(K), rng.randint(K)
        ref = np.dot(dy[:, o].reshape(-1).astype(np.float64), np.ascontiguousarray(xp[:, i, j:j + H, k:k + H]).reshape(-1).astype(np.float64))
        assert dwh[o, i, j, k] == np.float32(ref), (o, i, j, k, dwh[o, i, j, k], ref)
    exact("db", dbh, dy.sum((0, 2, 3), dtype=np.float64).astype(np.float32))
    # (3) whole tensors against the exact FFMA path (itself pinned to the oracle and the goldens in test_gpu_parity.py)
    if mode == "bf16":
        with cp.compute_mode("fp32_simt"):
            c = FunctionCache()
            y32 = Conv2DFn.forward(c, xt, wt, bt, 1, 1, 1)
            dx32, dw32, db32 = Conv2DFn.backward(c, dyt)
        exact("y vs fp32 path", yh, y32.to_numpy()); exact("dx vs fp32 path", dxh, dx32.to_numpy())
        exact("dw vs fp32 path", dwh, dw32.to_numpy()); exact("db vs fp32 path", dbh, db32.to_numpy())


@pytest.mark.parametrize("mode,tol", [("bf16", 1e-2), ("tf32", 2e-3), ("fp32", 1e-5), ("fp32_simt", 1e-5)])
@pytest.mark.parametrize("C", [128, 256, 512])
def test_conv2d_sweep_real_valued(cp, C, mode, tol):
    """(fp32_simt: C = 128 only — the FFMA kernels need seconds per pass beyond.)  Same shapes with real-valued data at B = 32 (enough pixels for every tile shape of the full-size run): complete
    planes of two images and sampled dw entries against fp64; tolerance = the mode's stated bound (bench.py TOL)."""
    import torch
    from compyute_b200.nn.functional import Conv2DFn, FunctionCache
    if mode == "fp32_simt" and C > 128:
        pytest.skip("FFMA path: covered at C = 128")
    B, H, K = 32, 56, 3
    rng = np.random.RandomState(7 + C)
    x = rng.uniform(-0.1, 0.1, (B, C, H, H)).astype(np.float32)
    kk = 1.0 / np.sqrt(C * K * K)
    w = rng.uniform(-kk, kk, (C, C, K, K)).astype(np.float32)
    b = rng.uniform(-kk, kk, (C,)).astype(np.float32)
    dy = rng.uniform(-0.1, 0.1, (B, C, H, H)).astype(np.float32)
    T = lambda a: cp.tensor(a, device=cp.cuda)
    with cp.compute_mode(mode):
        c = FunctionCache()
        y = Conv2DFn.forward(c, T(x), T(w), T(b), 1, 1, 1)
        dx, dw, db = Conv2DFn.backward(c, T(dy))
    tc_ok()
    sel = [0, B - 1]
    ws = torch.from_numpy(w.astype(np.float64))
    y_ref = torch.nn.functional.conv2d(torch.from_numpy(x[sel].astype(np.float64)), ws, torch.from_numpy(b.astype(np.float64)), padding=1).numpy()
    dx_ref = torch.nn.functional.conv_transpose2d(torch.from_numpy(dy[sel].astype(np.float64)), ws, padding=1).numpy()
    yh, dxh, dwh = y.to_numpy(), dx.to_numpy(), dw.to_numpy()

    def within(name, got, ref):
        if mode in ("fp32", "fp32_simt"):
            assert np.allclose(got, ref, rtol=tol, atol=tol), f"{name}: max abs err {np.abs(got - ref).max():.3e}"
        else:
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err <= tol, f"{name}: rel-max err {err:.3e} > {tol}"

    within("y", yh[sel], y_ref); within("dx", dxh[sel], dx_ref)
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1))).astype(np.float64)
    idx = [(rng.randint(C), rng.randint(C), rng.randint(K), rng.randint(K)) for _ in range(32)]
    ref = np.array([np.dot(dy[:, o].reshape(-1).astype(np.float64), np.ascontiguousarray(xp[:, i, j:j + H, k:k + H]).reshape(-1)) for o, i, j, k in idx])
    got = np.array([dwh[t] for t in idx])
    scale = np.abs(dwh).max()
    if mode in ("fp32", "fp32_simt"):
        assert np.allclose(got, ref, rtol=2e-5, atol=2e-5 * scale), np.abs(got - ref).max()
    else:
        assert np.abs(got - ref).max() <= tol * scale
    assert np.allclose(db.to_numpy(), dy.sum((0, 2, 3), dtype=np.float64), rtol=1e-4, atol=1e-4)
