"""Round-2 host-side features on the device (``-m gpu``): foreign device memory adoption (``__cuda_array_interface__``),
cross-entropy over N-D logits, the device-side Dataloader gather, CUDA-graph capture without warm-up side effects, and the
checkpoint size of a model whose gradients live in an optimizer arena."""
import pickle

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


def test_cuda_array_interface_is_adopted_without_a_copy(cp):
    """``cp.tensor(obj)`` on anything exporting ``__cuda_array_interface__`` (CuPy in the reference, backend.py:169-173; a
    torch tensor and a bare exporter object here): same device pointer, results of the hot path identical to a copied-in tensor."""
    import torch
    from compyute_b200.nn.functional import conv2d
    rng = np.random.RandomState(0)
    x = rng.uniform(-1, 1, (2, 8, 10, 10)).astype(np.float32)
    w = rng.uniform(-1, 1, (4, 8, 3, 3)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()

    class Exporter:  # what a CuPy / Numba array looks like to a consumer
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = {"shape": tuple(t.shape), "typestr": "<f4", "data": (t.data_ptr(), False), "version": 3, "strides": None}

    for obj in (xt, Exporter(xt)):
        a = cp.tensor(obj)
        assert a.device == cp.cuda and a.shape == x.shape and a.data.ptr == xt.data_ptr()
        y = conv2d(a, cp.tensor(w, device=cp.cuda), None, 1, 1, 1)
        y_ref = conv2d(cp.tensor(x, device=cp.cuda), cp.tensor(w, device=cp.cuda), None, 1, 1, 1)
        assert np.array_equal(y.to_numpy(), y_ref.to_numpy())
    xt.mul_(2.0)  # shared memory: the adopted tensor sees the producer's write
    torch.cuda.synchronize()
    assert np.array_equal(cp.tensor(xt).to_numpy(), 2.0 * x)
    strided = type("StridedExporter", (), {"__cuda_array_interface__": {
        "shape": (2, 8), "typestr": "<f4", "data": (xt.data_ptr(), False), "version": 3, "strides": (4, 8)}})()
    with pytest.raises(cp.ShapeError):  # non-contiguous views are refused, not silently copied
        cp.tensor(strided)


@pytest.mark.parametrize("shape", [(8, 16), (8, 16, 32), (4, 3, 5, 12)], ids=str)
def test_cross_entropy_nd_logits(cp, shape):
    """loss_funcs.py:57-69 on (..., classes) logits: softmax over the last dim, mean over all leading dims."""
    from compyute_b200 import nn
    rng = np.random.RandomState(1)
    x = rng.normal(0, 1, shape).astype(np.float32)
    t = rng.randint(0, shape[-1], shape[:-1])
    rc = []
    loss_ref = R.cross_entropy_forward(rc, x, t)
    dx_ref = R.cross_entropy_backward(rc)
    loss_fn = nn.CrossEntropyLoss()
    loss = loss_fn(cp.tensor(x, device=cp.cuda), cp.tensor(t.astype(np.int64), device=cp.cuda))
    dx = loss_fn.backward()
    assert abs(loss.item() - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert dx.shape == shape and np.allclose(dx.to_numpy(), dx_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("nc", [130, 512, 1000, 4096, 10], ids=str)
def test_softmax_ce_register_kernels(cp, nc):
    """Every class-count branch of cpt_softmax_ce_fwd (warp-per-row registers, block-per-row registers, generic) vs the oracle."""
    from compyute_b200 import nn
    rng = np.random.RandomState(nc)
    x = (rng.normal(0, 3, (37, nc))).astype(np.float32)
    t = rng.randint(0, nc, (37,))
    rc = []
    loss_ref = R.cross_entropy_forward(rc, x, t)
    dx_ref = R.cross_entropy_backward(rc)
    loss_fn = nn.CrossEntropyLoss()
    loss = loss_fn(cp.tensor(x, device=cp.cuda), cp.tensor(t.astype(np.int32), device=cp.cuda))
    dx = loss_fn.backward()
    assert abs(loss.item() - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert np.allclose(dx.to_numpy(), dx_ref, rtol=1e-5, atol=1e-7)


def test_dataloader_device_resident_gather(cp):
    """Data already on cuda: batches are gathered on the device in the reference's shuffle order (numpy's legacy stream,
    dataloaders.py:62-66) and equal host slicing bit for bit; labels arrive as int32."""
    from compyute_b200.nn.utils import Dataloader
    rng = np.random.RandomState(3)
    X = rng.normal(0, 1, (103, 3, 6, 6)).astype(np.float32)
    T = rng.randint(0, 10, (103,)).astype(np.int64)
    dl = Dataloader((cp.tensor(X, device=cp.cuda), cp.tensor(T, device=cp.cuda)), batch_size=16, device=cp.cuda, shuffle_data=True)
    np.random.seed(5)
    got = [(x.to_numpy(), t.to_numpy()) for x, t in dl()]
    np.random.seed(5)
    idx = np.random.permutation(103)
    assert len(got) == len(dl) == 7
    for i, (x, t) in enumerate(got):
        b = idx[i * 16:(i + 1) * 16]
        assert np.array_equal(x, X[b]) and np.array_equal(t, T[b].astype(np.int32)) and t.dtype == np.int32
    # host-resident data through the pinned double buffer gives the same batches
    dl2 = Dataloader((cp.tensor(X), cp.tensor(T)), batch_size=16, device=cp.cuda, shuffle_data=True)
    np.random.seed(5)
    for (x, t), (x2, t2) in zip(got, dl2()):
        assert np.array_equal(x, x2.to_numpy()) and np.array_equal(t, t2.to_numpy())


def test_captured_step_restore_after_warmup(cp):
    """``CapturedStep(..., restore_after_warmup=True)``: the eager warm-up leaves parameters, running statistics and optimizer
    state (moments, t) untouched, so the first replay is step 1 — identical to an eager run from the same start."""
    import bench_workloads as W
    from compyute_b200 import nn
    spec = W.mnist_cnn(drop=0.0)
    rng = np.random.RandomState(9)
    xs = [rng.normal(0, 1, (16, 1, 28, 28)).astype(np.float32) for _ in range(3)]
    ts = [rng.randint(0, 10, (16,)).astype(np.int32) for _ in range(3)]

    def make():
        np.random.seed(21)
        with cp.use_device(cp.cuda):
            m = W.build(spec)
        m.training()
        return m, nn.optimizers.Adam(m.get_parameters(), lr=1e-3), nn.CrossEntropyLoss()

    m1, o1, l1 = make()
    eager = []
    for x, t in zip(xs, ts):
        loss = l1(m1(cp.tensor(x, device=cp.cuda)), cp.tensor(t, device=cp.cuda))
        o1.reset_grads(); m1.backward(l1.backward()); o1.step()
        eager.append(loss.item())
    m2, o2, l2 = make()
    x_s, t_s = cp.tensor(xs[0], device=cp.cuda), cp.tensor(ts[0], device=cp.cuda)

    def step():
        loss = l2(m2(x_s), t_s)
        o2.reset_grads(); m2.backward(l2.backward()); o2.step()
        return loss

    cap = cp.graph.CapturedStep(step, optimizers=[o2], warmup=3, restore_after_warmup=True, modules=[m2])
    assert o2.t == 1
    got = []
    for x, t in zip(xs, ts):
        x_s.data.upload(x); t_s.data.upload(t)
        got.append(cap().item())
    assert got == eager, (got, eager)
    for a, b in zip(m1.get_state_dict().values(), m2.get_state_dict().values()):
        assert np.array_equal(a.to_numpy(), b.to_numpy())
    # a buffer rebound after the capture is detected instead of silently updating the orphaned array
    bn = next(m for m in m2.get_modules() if hasattr(m, "rmean"))
    bn.rmean.data = bn.rmean.data.copy()
    with pytest.raises(RuntimeError):
        cap()


def test_checkpoint_does_not_carry_gradient_arena(cp):
    """``cp.save(model.get_state_dict())`` pickles parameter VALUES only: not the gradient, not the optimizer's arena slot."""
    from compyute_b200 import nn
    np.random.seed(0)
    with cp.use_device(cp.cuda):
        m = nn.Sequential(nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 8))
    m.training()
    opt = nn.optimizers.SGD(m.get_parameters(), lr=0.1)
    x = cp.tensor(np.random.RandomState(0).normal(0, 1, (4, 64)).astype(np.float32), device=cp.cuda)
    m(x)
    m.backward(cp.tensor(np.ones((4, 8), np.float32), device=cp.cuda))
    assert all(p.grad is not None and p.grad_slot is not None for p in m.get_parameters())
    blob = pickle.dumps(m.get_state_dict())
    n_values = sum(p.size for p in m.get_parameters()) * 4
    assert len(blob) < 1.2 * n_values + 4096, (len(blob), n_values)
    sd = pickle.loads(blob)
    for k, v in sd.items():
        assert v.grad is None and getattr(v, "grad_slot", None) is None
        assert np.array_equal(v.to_numpy(), m.get_state_dict()[k].to_numpy())
    opt.step()


# ------------------------------------------------------------------ Conv2D -> ReLU from the convolution's epilogue (f4)
@pytest.mark.parametrize("mode", ["bf16", "tf32", "fp32"])
def test_conv_relu_epilogue_fusion_is_bit_exact(cp, mode):
    """Conv2D -> ReLU inside a Sequential: the ReLU comes out of the convolution's GEMM epilogue (cpt_conv2d_fprop_cl_relu /
    _packed_relu) and its backward is folded into the staging of dy (cpt_to_channels_last_gated, mask = fused output > 0).
    Outputs, input gradient and every parameter gradient must equal the layer-by-layer evaluation bit for bit — packed first
    layer (Ci = 1), ragged sizes, stride, dilation, with and without bias, NaN inputs, inference mode — and the fused walk
    launches fewer kernels.  A pair that ENDS a container (the block of a ResidualConnection) is left alone."""
    from compyute_b200 import _lib, nn
    from compyute_b200.nn.functional.activation_funcs import FUSED_INTO_PRODUCER
    rng = np.random.RandomState(17)
    for B, H, nan in ((5, 28, False), (3, 19, True)):
        x = rng.normal(0, 1, (B, 1, H, H)).astype(np.float32)
        if nan:
            x[1, 0, 4, 7] = np.nan

        def build():
            np.random.seed(23)
            with cp.use_device(cp.cuda):
                return nn.Sequential(
                    nn.Conv2D(1, 32, 5), nn.ReLU(),                                    # packed first layer (bf16 mode)
                    nn.Conv2D(32, 32, 5, bias=False), nn.ReLU(), nn.MaxPooling2D(1),
                    nn.Conv2D(32, 64, 3, stride=2, padding=1), nn.ReLU(),
                    nn.ResidualConnection(nn.Conv2D(64, 64, 3, padding="same"), nn.ReLU()),  # ends its container: not fused
                    nn.Conv2D(64, 40, 3, dilation=2, padding="same"), nn.ReLU(),
                    nn.Conv2D(40, 8, 3, padding="same"))

        def run(fused):
            nn.set_fusion_enabled(fused)
            try:
                model = build()
                model.training()
                n0 = _lib.lib().cpt_launch_count()
                with cp.compute_mode(mode):
                    y = model(cp.tensor(x, device=cp.cuda))
                    marks = sum(1 for m in model.get_modules() if type(m) is nn.ReLU and m.fcache.cache
                                and m.fcache.cache[-1][0] is FUSED_INTO_PRODUCER)
                    dy = np.random.RandomState(6).normal(0, 1, y.shape).astype(np.float32)
                    dx = model.backward(cp.tensor(dy, device=cp.cuda))
                    launches = _lib.lib().cpt_launch_count() - n0
                    assert _lib.lib().cpt_tc_check_status() == 0
                    assert all(not m.fcache.cache for m in model.get_modules())
                    out = [y.to_numpy(), dx.to_numpy()] + [p.grad.to_numpy() for p in model.get_parameters()]
                    model.inference()
                    out.append(model(cp.tensor(x, device=cp.cuda)).to_numpy())
                return out, launches, marks
            finally:
                nn.set_fusion_enabled(True)

        (a, la, ma), (b, lb, mb) = run(True), run(False)
        # fp32 mode: the Ci = 1 layer runs on the exact FFMA kernels (not fusable), the other three pairs are
        assert ma == (3 if mode == "fp32" else 4) and mb == 0, (ma, mb)
        for k, (u, v) in enumerate(zip(a, b)):
            assert np.array_equal(u, v, equal_nan=True), (mode, H, k, u.shape, float(np.nanmax(np.abs(u - v))))
        assert la < lb, (la, lb)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_relu_function_vs_oracle(cp, mode):
    """``Conv2DFn.forward(..., relu=True)`` / ``backward`` against the oracle's Conv2D and ReLU functions chained
    (convolution_funcs.py:218-254, activation_funcs.py:26-34): 1e-5 in the exact mode, the bf16 tolerance otherwise."""
    from compyute_b200.nn.functional import Conv2DFn, FunctionCache
    rng = np.random.RandomState(3)
    for (B, Ci, Co, H, K, P, s, d, bias) in ((4, 16, 24, 14, 3, 1, 1, 1, True), (2, 40, 72, 17, 3, 2, 2, 1, False), (3, 8, 130, 9, 5, 4, 1, 2, True)):
        x = rng.uniform(-1, 1, (B, Ci, H, H)).astype(np.float32)
        w = (rng.uniform(-1, 1, (Co, Ci, K, K)) / np.sqrt(Ci * K * K)).astype(np.float32)
        b = rng.uniform(-0.3, 0.3, (Co,)).astype(np.float32) if bias else None
        rc, rr = [], []
        a_ref = R.conv2d_forward(rc, x, w, b, P, s, d)
        y_ref = R.relu_forward(rr, a_ref)
        dy = rng.uniform(-1, 1, y_ref.shape).astype(np.float32)
        dx_ref, dw_ref, db_ref = R.conv2d_backward(rc, R.relu_backward(rr, dy))
        T = lambda t: cp.tensor(t, device=cp.cuda)
        with cp.compute_mode(mode):
            assert Conv2DFn.relu_fusable(T(x), T(w), P, s, d)
            c = FunctionCache()
            y = Conv2DFn.forward(c, T(x), T(w), None if b is None else T(b), P, s, d, False, relu=True)
            dx, dw, db = Conv2DFn.backward(c, T(dy))
        yv = y.to_numpy()
        # an activation within rounding distance of 0 may land on the other side of the ReLU: the forward is compared where |a|
        # is clear of it, the backward against the oracle evaluated with the mask the device produced
        clear = np.abs(a_ref) > (1e-5 if mode == "fp32" else 2e-2)
        tol = 1e-5 if mode == "fp32" else 1e-2
        assert np.allclose(yv[clear], y_ref[clear], rtol=tol, atol=tol * max(1.0, float(np.abs(y_ref).max())))
        assert ((yv > 0) == (a_ref > 0))[clear].all()
        rc2 = []
        R.conv2d_forward(rc2, x, w, b, P, s, d)
        dx_ref, dw_ref, db_ref = R.conv2d_backward(rc2, dy * (yv > 0))
        for got, ref, t32 in ((dx, dx_ref, 1e-5), (dw, dw_ref, 2e-5)) + (((db, db_ref, 1e-4),) if bias else ()):
            if mode == "fp32":
                assert np.allclose(got.to_numpy(), ref, rtol=t32, atol=t32), float(np.abs(got.to_numpy() - ref).max())
            else:
                assert np.abs(got.to_numpy() - ref).max() <= 1e-2 * max(float(np.abs(ref).max()), 1e-30)
        assert (y.to_numpy() >= 0).all()
