"""Model-level GPU parity (``-m gpu``): reduced-size instances of BASELINE.json's configs 1, 3, 4, 5 trained for a few
steps through the compyute_b200 module API, against the oracle's model interpreter (same spec, same initial parameters,
same batch).  fp32 mode.  Checked: every parameter gradient of the first backward (error <= 2e-4 of the gradient's max
magnitude: per-op agreement is 1e-5, test_gpu_parity.py, and a 20-60-op deep fp32 network compounds it), the loss trace
(1e-4) and the parameters after 3 SGD-momentum steps.  SGD rather than Adam on purpose: several parameters of these models
have an analytically ZERO gradient (a conv bias feeding a BatchNorm, possibly through an all-positive ReLU channel), which
both implementations return as rounding noise of different sign, and Adam's g/sqrt(g^2) turns that noise into +-lr."""
import numpy as np
import pytest

import bench_workloads as W
from oracle.model_ref import RefModel

pytestmark = pytest.mark.gpu

CASES = {
    "c1_mnist_cnn": (W.mnist_cnn(drop=0.0), (8, 1, 28, 28), 10),
    "c3_vgg_w8": (W.vgg(width=8, hw=32, hidden=32), (6, 3, 32, 32), 10),
    "c4_resnet18_w8": (W.resnet18(width=8, hw=64, classes=16), (4, 3, 64, 64), 16),
    "c5_mlp_w64": (W.mlp(width=64, depth=8), (16, 64), 64),
}


@pytest.mark.parametrize("name", list(CASES), ids=list(CASES))
def test_config_train_steps(name):
    import compyute_b200 as cp
    from compyute_b200 import nn
    spec, xshape, classes = CASES[name]
    np.random.seed(11)
    with cp.use_device(cp.cuda):
        model = W.build(spec)
    model.training()
    params0 = [p.to_numpy().copy() for p in model.get_parameters()]
    bufs0 = [b.to_numpy().copy() for b in model.get_buffers()]
    rng = np.random.RandomState(5)
    x = rng.normal(0, 1, xshape).astype(np.float32)
    t = rng.randint(0, classes, (xshape[0],))
    ref = RefModel(spec, [p.copy() for p in params0], [b.copy() for b in bufs0])
    # gradients of the first backward pass
    from oracle import compyute_ref as R
    gref = RefModel(spec, [p.copy() for p in params0], [b.copy() for b in bufs0])
    lc = []
    R.cross_entropy_forward(lc, gref.forward(x, True), t)
    gref.backward(R.cross_entropy_backward(lc))
    ref_losses = ref.train_steps(x, t, 3, lr=0.05, optimizer="sgd", momentum=0.9)

    opt = nn.optimizers.SGD(model.get_parameters(), lr=0.05, momentum=0.9)
    loss_fn = nn.CrossEntropyLoss()
    xt, tt = cp.tensor(x, device=cp.cuda), cp.tensor(t.astype(np.int32), device=cp.cuda)
    losses = []
    for step in range(3):
        loss = loss_fn(model(xt), tt)
        opt.reset_grads()
        model.backward(loss_fn.backward())
        if step == 0:
            for i, (p, g) in enumerate(zip(model.get_parameters(), gref.gradients())):
                got = p.grad.to_numpy()
                assert got.shape == g.shape
                # floor 1e-2: the analytically ZERO gradients (conv bias feeding a BatchNorm) are ~1e-7 rounding noise on both sides
                assert np.abs(got - g).max() <= 2e-4 * max(np.abs(g).max(), 1e-2), f"grad {i}: {np.abs(got - g).max():.3e} vs max {np.abs(g).max():.3e}"
        opt.step()
        losses.append(loss.item())
    assert np.allclose(losses, ref_losses, rtol=1e-4, atol=1e-5), (losses, ref_losses)
    for i, (p, r) in enumerate(zip(model.get_parameters(), ref.parameters())):
        assert p.shape == r.shape
        assert np.allclose(p.to_numpy(), r, rtol=5e-4, atol=5e-5), f"param {i} max err {np.abs(p.to_numpy() - r).max():.3e}"
    for b, r in zip(model.get_buffers(), ref.buffers()):
        assert np.allclose(b.to_numpy(), r, rtol=1e-4, atol=1e-5)
    assert all(not m.fcache.cache for m in model.get_modules())


# Tensor-core modes against the ORACLE (not against the repo's own fp32 run).  Widths are chosen so that the layers run on the
# kernels the benchmarks use (im2col tcgen05 path for Ci >= 32, packed-K path for the first layer, fused Linear+ReLU
# epilogues for the MLP).  Two oracles, two stated tolerances per case and mode:
#   (A) "rounded": the oracle with the mode's OPERAND ROUNDING emulated (RefModel(operand_rounding=mode): x, w, dy rounded to
#       bf16 / tf32 before each fp32 contraction — tools/rounding_probe.py pins the emulation per operation to ~1e-6 on the
#       device: bf16 = RNE, tf32 convolutions = RNA in the staging kernels, tf32 Linear = the tensor core's truncation).
#       For the MLP this reproduces the device run to rounding noise (loss bit-equal, gradients ~1e-6).  For the CNNs it
#       cannot stay that tight, and the bound says by how much: rounding is DISCONTINUOUS.  An activation that differs from
#       the oracle's by 1e-7 (fp32 summation order) rounds to the neighbouring bf16 value with probability |delta|/ulp; one
#       such flip moves all 9*C outputs of the next convolution that read it by ulp*|w| ~ 2e-4 of their scale, which makes
#       more inputs of the following layer flip, and BatchNorm over a batch of 4 (all the einsum oracle affords) divides by
#       small standard deviations on the way.  tools/model_parity_probe.py shows the layer-by-layer growth (profiles/
#       r02_model_parity_probe.log): 2e-7 after the first conv/BN/ReLU, 2e-4 after the second conv, 5e-3 at the logits (bf16).
#   (B) "fp32": the reference's own arithmetic.  The loss agrees to the mode's operand precision; parameter gradients
#       additionally see every ReLU / max-pool decision within one operand-rounding of a tie flipped (a flipped unit changes
#       its whole gradient contribution; a fraction f of flipped units moves the gradient by ~sqrt(f) in relative L2).
# Bounds are the measured values (B200, round 2, gpurun_out/model_parity_tc_modes.jsonl -> profiles/) with >= 2x margin.
# Gradients that are analytically zero (a conv bias feeding a BatchNorm) are skipped: both sides hold rounding noise.
TC_CASES = {
    "c3_vgg_w32": (W.vgg(width=32, hw=32, hidden=64), (4, 3, 32, 32), 10),
    "c4_resnet18_w32": (W.resnet18(width=32, hw=64, classes=16), (4, 3, 64, 64), 16),
    "c5_mlp_w256": (W.mlp(width=256, depth=8), (64, 256), 256),
}
_ORACLE_RUNS = {}
TC_TOL_ROUNDED = {
    ("c5_mlp_w256", "bf16"): dict(loss=1e-5, grad=5e-3, trace=1e-4), ("c5_mlp_w256", "tf32"): dict(loss=1e-5, grad=2e-2, trace=1e-4),
    ("c3_vgg_w32", "bf16"): dict(loss=6e-4, grad=0.4, trace=2e-2), ("c3_vgg_w32", "tf32"): dict(loss=1e-4, grad=6e-2, trace=5e-3),
    ("c4_resnet18_w32", "bf16"): dict(loss=1.5e-3, grad=0.5, trace=5e-2), ("c4_resnet18_w32", "tf32"): dict(loss=6e-4, grad=0.3, trace=6e-2),
}
TC_TOL_FP32 = {"bf16": dict(loss=1e-2, grad=0.8, trace=0.1), "tf32": dict(loss=2e-3, grad=0.3, trace=3e-2)}


def _oracle_run(spec, params0, bufs0, x, t, rounding):
    from oracle import compyute_ref as R
    gref = RefModel(spec, [p.copy() for p in params0], [b.copy() for b in bufs0], operand_rounding=rounding)
    lc = []
    loss0 = float(R.cross_entropy_forward(lc, gref.forward(x, True), t))
    gref.backward(R.cross_entropy_backward(lc))
    ref = RefModel(spec, [p.copy() for p in params0], [b.copy() for b in bufs0], operand_rounding=rounding)
    return gref.gradients(), loss0, ref.train_steps(x, t, 3, lr=0.02, optimizer="sgd", momentum=0.9)


@pytest.mark.parametrize("mode", ["bf16", "tf32"])
@pytest.mark.parametrize("name", list(TC_CASES), ids=list(TC_CASES))
def test_config_train_steps_tensor_core_modes_vs_oracle(name, mode):
    import json
    import os

    import compyute_b200 as cp
    from compyute_b200 import _lib, nn
    spec, xshape, classes = TC_CASES[name]
    np.random.seed(11)
    with cp.use_device(cp.cuda):
        model = W.build(spec)
    model.training()
    params0 = [p.to_numpy().copy() for p in model.get_parameters()]
    bufs0 = [b.to_numpy().copy() for b in model.get_buffers()]
    rng = np.random.RandomState(5)
    x = rng.normal(0, 1, xshape).astype(np.float32)
    t = rng.randint(0, classes, (xshape[0],))
    for rounding in (mode, None):  # oracle runs cost seconds per step on the einsum path: the fp32 one is shared by the modes
        if (name, rounding) not in _ORACLE_RUNS:
            _ORACLE_RUNS[(name, rounding)] = _oracle_run(spec, params0, bufs0, x, t, rounding)

    opt = nn.optimizers.SGD(model.get_parameters(), lr=0.02, momentum=0.9)
    loss_fn = nn.CrossEntropyLoss()
    xt, tt = cp.tensor(x, device=cp.cuda), cp.tensor(t.astype(np.int32), device=cp.cuda)
    losses, grads = [], None
    with cp.compute_mode(mode):
        for step in range(3):
            loss = loss_fn(model(xt), tt)
            opt.reset_grads()
            model.backward(loss_fn.backward())
            if step == 0:
                grads = [p.grad.to_numpy().astype(np.float64) for p in model.get_parameters()]
            opt.step()
            losses.append(loss.item())
    assert _lib.lib().cpt_tc_check_status() == 0
    report = {"case": name, "mode": mode}
    failures = []
    for tag, rounding, tol in (("rounded_oracle", mode, TC_TOL_ROUNDED[(name, mode)]), ("fp32_oracle", None, TC_TOL_FP32[mode])):
        g_ref, loss0, ref_losses = _ORACLE_RUNS[(name, rounding)]
        gmax = max(np.abs(g).max() for g in g_ref)
        gerrs = [(i, float(np.linalg.norm(got - g) / np.linalg.norm(g))) for i, (got, g) in enumerate(zip(grads, g_ref))
                 if np.abs(g).max() > 1e-5 * gmax]
        loss_err = abs(losses[0] - loss0) / max(1.0, abs(loss0))
        trace_err = float(np.max(np.abs(np.array(losses) - np.array(ref_losses)) / np.maximum(1.0, np.abs(ref_losses))))
        worst = max(gerrs, key=lambda e: e[1])
        report[tag] = {"loss_err": loss_err, "trace_err": trace_err, "worst_grad": worst, "median_grad": float(np.median([e for _, e in gerrs])), "tol": tol}
        if loss_err > tol["loss"]:
            failures.append(f"{tag}: first loss {losses[0]} vs {loss0}")
        if worst[1] > tol["grad"]:
            failures.append(f"{tag}: parameter {worst[0]} relative L2 gradient error {worst[1]:.3e} > {tol['grad']}")
        if trace_err > tol["trace"]:
            failures.append(f"{tag}: loss trace {losses} vs {ref_losses}")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):  # measured margins, kept next to the run's other artefacts
        with open(os.path.join(out, "model_parity_tc_modes.jsonl"), "a") as f:
            f.write(json.dumps(report) + "\n")
    assert not failures, failures


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_vgg_tensor_core_modes_track_fp32(mode):
    """Same VGG-style model in a tensor-core mode: the first-step loss matches fp32 within the mode's tolerance and
    training still decreases the loss."""
    import compyute_b200 as cp
    from compyute_b200 import nn
    spec = W.vgg(width=16, hw=32, hidden=64)
    rng = np.random.RandomState(2)
    x = rng.normal(0, 1, (32, 3, 32, 32)).astype(np.float32)
    t = rng.randint(0, 10, (32,)).astype(np.int32)

    def run(m):
        np.random.seed(3)
        with cp.use_device(cp.cuda):
            model = W.build(spec)
        model.training()
        opt = nn.optimizers.Adam(model.get_parameters(), lr=2e-3)
        loss_fn = nn.CrossEntropyLoss()
        xt, tt = cp.tensor(x, device=cp.cuda), cp.tensor(t, device=cp.cuda)
        out = []
        with cp.compute_mode(m):
            for _ in range(8):
                loss = loss_fn(model(xt), tt)
                opt.reset_grads(); model.backward(loss_fn.backward()); opt.step()
                out.append(loss.item())
        return out

    ref, got = run("fp32"), run(mode)
    assert abs(got[0] - ref[0]) <= {"tf32": 2e-3, "bf16": 1e-2}[mode] * max(1.0, abs(ref[0]))
    assert got[-1] < got[0] and np.isfinite(got).all()


def test_cuda_graph_captured_step_matches_eager():
    """A train step captured into a CUDA graph (cp.graph.CapturedStep) replays to exactly the eager results: same losses,
    bit-identical parameters, BatchNorm running stats and Adam state after 6 steps, with lr changed between replays."""
    import compyute_b200 as cp
    from compyute_b200 import nn
    spec = W.mnist_cnn(drop=0.0)
    rng = np.random.RandomState(9)
    xs = [rng.normal(0, 1, (32, 1, 28, 28)).astype(np.float32) for _ in range(6)]
    ts = [rng.randint(0, 10, (32,)).astype(np.int32) for _ in range(6)]

    def make():
        np.random.seed(21)
        with cp.use_device(cp.cuda):
            m = W.build(spec)
        m.training()
        return m, nn.optimizers.Adam(m.get_parameters(), lr=1e-3), nn.CrossEntropyLoss()

    def lr_at(i):
        return 1e-3 * (0.5 if i >= 5 else 1.0)

    # eager
    m1, o1, l1 = make()
    eager = []
    for i in range(6):
        o1.lr = lr_at(i)
        loss = l1(m1(cp.tensor(xs[i], device=cp.cuda)), cp.tensor(ts[i], device=cp.cuda))
        o1.reset_grads(); m1.backward(l1.backward()); o1.step()
        eager.append(loss.item())
    # captured: 3 eager warm-up steps inside CapturedStep (on batches 0..2), then replays for batches 3..5
    m2, o2, l2 = make()
    xs_t, ts_t = cp.tensor(xs[0], device=cp.cuda), cp.tensor(ts[0], device=cp.cuda)
    feed = iter(range(6))

    def step():
        loss = l2(m2(xs_t), ts_t)
        o2.reset_grads(); m2.backward(l2.backward()); o2.step()
        return loss

    class Feeder:  # warm-up calls inside CapturedStep consume batches 0, 1, 2
        n = 0
    def step_with_feed():
        i = Feeder.n
        if not cp.graph.is_capturing():
            xs_t.data.upload(xs[i]); ts_t.data.upload(ts[i]); o2.lr = lr_at(i); Feeder.n += 1
        return step()

    captured = cp.graph.CapturedStep(step_with_feed, optimizers=[o2], warmup=3)
    got = []
    for i in range(3, 6):
        xs_t.data.upload(xs[i]); ts_t.data.upload(ts[i]); o2.lr = lr_at(i)
        got.append(captured().item())
    assert got == eager[3:], (got, eager[3:])  # every kernel on the path reduces in a fixed order: bit-identical
    assert o2.t == o1.t == 7
    for a, b in zip(m1.get_state_dict().values(), m2.get_state_dict().values()):
        assert np.array_equal(a.to_numpy(), b.to_numpy())
    for i in o1._state:
        assert np.array_equal(o1._state[i]["v"].to_numpy(), o2._state[i]["v"].to_numpy())
