"""CPU test of the oracle's model interpreter (oracle/model_ref.py): analytic gradients of a small residual CNN agree with
central finite differences in fp64, and the FLOP accounting of bench_workloads matches SURVEY §8d."""
import numpy as np

import bench_workloads as W
from oracle import compyute_ref as R
from oracle.model_ref import RefModel

SPEC = [("conv", 2, 4, 3, 1, 1, True), ("bn2d", 4), ("relu",),
        ("residual", [("conv", 4, 4, 3, 1, 1, False), ("bn2d", 4), ("relu",), ("conv", 4, 4, 3, 1, 1, False), ("bn2d", 4)], None), ("relu",),
        ("residual", [("conv", 4, 6, 3, 1, 2, False), ("bn2d", 6)], [("conv", 4, 6, 1, 0, 2, False), ("bn2d", 6)]), ("relu",),
        ("maxpool", 2), ("flatten",), ("linear", 6 * 2 * 2, 7, True), ("bn1d", 7), ("relu",), ("linear", 7, 5, True)]


def init(spec, rng):
    ps, bs = [], []
    for s in spec:
        if s[0] == "conv":
            ps.append(rng.normal(0, 0.3, (s[2], s[1], s[3], s[3])))
            if s[6]:
                ps.append(rng.normal(0, 0.1, (s[2],)))
        elif s[0] == "linear":
            ps.append(rng.normal(0, 0.3, (s[2], s[1])))
            if s[3]:
                ps.append(rng.normal(0, 0.1, (s[2],)))
        elif s[0] in ("bn2d", "bn1d"):
            ps += [rng.uniform(0.5, 1.5, s[1]), rng.normal(0, 0.1, s[1])]
            bs += [np.zeros(s[1]), np.ones(s[1])]
        elif s[0] == "residual":
            for sub in (s[1], s[2] or []):
                p2, b2 = init(sub, rng)
                ps += p2; bs += b2
    return ps, bs


def test_model_gradients_match_finite_differences():
    rng = np.random.RandomState(0)
    ps, bs = init(SPEC, rng)
    x = rng.normal(0, 1, (4, 2, 8, 8)); t = rng.randint(0, 5, (4,))

    def loss_of(params):
        m = RefModel(SPEC, [p.copy() for p in params], [b.copy() for b in bs]); lc = []
        return float(R.cross_entropy_forward(lc, m.forward(x, True), t))

    m = RefModel(SPEC, [p.copy() for p in ps], [b.copy() for b in bs]); lc = []
    R.cross_entropy_forward(lc, m.forward(x, True), t)
    m.backward(R.cross_entropy_backward(lc))
    g = m.gradients()
    assert len(g) == len(ps) == len(m.parameters())
    for pi in range(len(ps)):
        for _ in range(2):
            idx = tuple(rng.randint(0, s) for s in ps[pi].shape); e = 1e-5
            pp = [p.copy() for p in ps]
            pp[pi][idx] += e; lp = loss_of(pp)
            pp[pi][idx] -= 2 * e; lm = loss_of(pp)
            fd = (lp - lm) / (2 * e)
            assert abs(fd - g[pi][idx]) <= 2e-3 * abs(fd) + 1e-7, (pi, idx, fd, g[pi][idx])
    # training mode updates the running stats, inference mode does not
    b0 = [b.copy() for b in m.buffers()]
    m.forward(x, False)
    assert all(np.array_equal(a, b) for a, b in zip(b0, m.buffers()))


def test_flop_accounting_matches_survey():
    assert abs(W.train_flops_per_image(W.mnist_cnn(), 28) / 80.4e6 - 1) < 0.01
    assert abs(W.train_flops_per_image(W.vgg(), 32) / 1.26e9 - 1) < 0.01
    assert abs(W.train_flops_per_image(W.resnet18(), 224) / 10.9e9 - 1) < 0.01
    assert abs(W.train_flops_per_image(W.mlp(), 1) * 8192 / 6.60e12 - 1) < 0.01
