"""Generate golden vectors from the REAL reference (Compyute v0.1.8 at /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference is not on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference hard-imports ``cupy`` (compyute/backend.py:12) and ``tensorboardX``
(compyute/nn/utils/tensorboard.py:3), neither installed here, so tiny import stubs are written to a
temporary directory first.  Only the reference's NumPy path (device=cpu) is executed.
Fixtures follow the reference's own tests: legacy ``numpy.random.seed(42)`` + ``uniform(-0.1, 0.1)``
for activations, ``uniform(-1, 1) * 0.1`` for parameters (tests/utils.py:15-51).
"""

from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference"


def _install_shims() -> None:
    d = tempfile.mkdtemp(prefix="cpt_shim_")
    os.makedirs(os.path.join(d, "cupy"))
    with open(os.path.join(d, "cupy", "__init__.py"), "w") as f:
        f.write(
            "import numpy as _np\n"
            "class ndarray: pass\n"
            "def asnumpy(a): return _np.asarray(a)\n"
            "def asarray(a): return _np.asarray(a)\n"
            "class _Dev:\n"
            "    def __init__(self, *a): pass\n"
            "class cuda:\n"
            "    Device = _Dev\n"
            "    @staticmethod\n"
            "    def is_available(): return False\n"
            "class random:\n"
            "    @staticmethod\n"
            "    def seed(*a): pass\n"
        )
    os.makedirs(os.path.join(d, "tensorboardX"))
    with open(os.path.join(d, "tensorboardX", "__init__.py"), "w") as f:
        f.write("class SummaryWriter: pass\n")
    sys.path.insert(0, d)
    sys.path.insert(0, REF)


def rnd(shape, lo=-0.1, hi=0.1, seed=42):
    np.random.seed(seed)
    return np.random.uniform(lo, hi, shape).astype(np.float32)


def main() -> None:
    _install_shims()
    import compyute as cp  # the reference
    from compyute.nn.functional.activation_funcs import ReLUFn
    from compyute.nn.functional.convolution_funcs import Conv2DFn
    from compyute.nn.functional.functions import FunctionCache
    from compyute.nn.functional.linear_funcs import LinearFn
    from compyute.nn.functional.loss_funcs import CrossEntropyLossFn
    from compyute.nn.functional.normalization_funcs import BatchNorm1DFn, BatchNorm2DFn
    from compyute.nn.functional.pooling_funcs import AvgPooling2DFn, MaxPooling2DFn
    from compyute.nn.functional.regularization_funcs import DropoutFn
    from compyute.nn.modules.convolutions import _str_to_pad
    from compyute.nn.parameter import Parameter

    os.makedirs(OUT, exist_ok=True)
    T = cp.tensor
    manifest: dict[str, list] = {}

    def npy(t):
        return None if t is None else t.to_numpy()

    # ---------------- conv2d: the 12 combinations of tests/nn/modules/convolutions_test.py:29-42,
    # at reduced sizes (the einsum path is ~0.5 GFLOP/s) + bias-free + 1x1/stride-2 + 7x7/s2/p3 cases
    conv = {}
    cases = []
    for (B, Ci, Co, H, K) in [(3, 3, 5, 12, 5), (2, 1, 6, 11, 3)]:
        for pad, s, d in [("valid", 1, 1), ("valid", 1, 2), ("valid", 2, 1), ("valid", 2, 2),
                          ("same", 1, 1), ("same", 1, 2)]:
            cases.append((B, Ci, Co, H, K, _str_to_pad(pad, K, d), s, d, True))
    cases += [(2, 4, 3, 9, 3, 1, 1, 1, False), (2, 4, 6, 8, 1, 0, 2, 1, True), (2, 3, 4, 14, 7, 3, 2, 1, True),
              (2, 8, 8, 6, 3, 1, 2, 1, False), (1, 2, 2, 5, 2, 1, 3, 1, True)]
    for n, (B, Ci, Co, H, K, P, s, d, bias) in enumerate(cases):
        x = rnd((B, Ci, H, H))
        w = rnd((Co, Ci, K, K), -1, 1) * np.float32(0.1)
        b = (rnd((Co,), -1, 1) * np.float32(0.1)) if bias else None
        c = FunctionCache()
        y = Conv2DFn.forward(c, T(x), T(w), None if b is None else T(b), P, s, d)
        dy = rnd(y.shape, seed=43)
        dx, dw, db = Conv2DFn.backward(c, T(dy))
        assert len(c.cache) == 0
        for k, v in dict(x=x, w=w, b=b, y=npy(y), dy=dy, dx=npy(dx), dw=npy(dw), db=npy(db)).items():
            if v is not None:
                conv[f"c{n}_{k}"] = v
        manifest.setdefault("conv2d", []).append(
            dict(id=n, B=B, Ci=Ci, Co=Co, H=H, K=K, padding=P, stride=s, dilation=d, bias=bias))
    np.savez_compressed(os.path.join(OUT, "conv2d.npz"), **conv)

    # ---------------- linear (tests/nn/modules/linear_test.py:10: 2-D / 3-D / 4-D)
    lin = {}
    for n, (shape, Cout, bias) in enumerate([((8, 16), 12, True), ((4, 5, 16), 12, True),
                                             ((2, 3, 4, 16), 7, True), ((8, 16), 12, False)]):
        x = rnd(shape)
        w = rnd((Cout, shape[-1]), -1, 1) * np.float32(0.1)
        b = (rnd((Cout,), -1, 1) * np.float32(0.1)) if bias else None
        c = FunctionCache()
        y = LinearFn.forward(c, T(x), T(w), None if b is None else T(b))
        dy = rnd(y.shape, seed=43)
        dx, dw, db = LinearFn.backward(c, T(dy))
        for k, v in dict(x=x, w=w, b=b, y=npy(y), dy=dy, dx=npy(dx), dw=npy(dw), db=npy(db)).items():
            if v is not None:
                lin[f"c{n}_{k}"] = v
        manifest.setdefault("linear", []).append(dict(id=n, shape=list(shape), Cout=Cout, bias=bias))
    np.savez_compressed(os.path.join(OUT, "linear.npz"), **lin)

    # ---------------- pooling (poolings_test.py:9-14: k in {2,3}; k=3 leaves an uncovered tail)
    pool = {}
    pcases = [((2, 3, 8, 8), 2, False), ((2, 3, 8, 8), 3, False), ((3, 2, 7, 7), 2, False),
              ((2, 2, 6, 6), 2, True), ((2, 2, 7, 7), 3, True), ((1, 1, 4, 4), 4, False)]
    for n, (shape, k, ties) in enumerate(pcases):
        if ties:  # few distinct values → many tied maxima, zeros in the tail, negative dy → -0.0
            np.random.seed(42)
            x = np.random.randint(0, 3, shape).astype(np.float32)
        else:
            x = rnd(shape)
        for name, Fn in (("max", MaxPooling2DFn), ("avg", AvgPooling2DFn)):
            c = FunctionCache()
            y = Fn.forward(c, T(x), k)
            dy = rnd(y.shape, seed=43)
            dx = Fn.backward(c, T(dy))
            pool.update({f"{name}{n}_x": x, f"{name}{n}_y": npy(y), f"{name}{n}_dy": dy, f"{name}{n}_dx": npy(dx)})
        manifest.setdefault("pool", []).append(dict(id=n, shape=list(shape), k=k, ties=ties))
    # NaN inside a window: y is NaN, mask all-False there (Appendix A.5)
    x = rnd((1, 1, 4, 4)); x[0, 0, 1, 1] = np.nan
    c = FunctionCache(); y = MaxPooling2DFn.forward(c, T(x), 2); dy = rnd(y.shape, seed=43)
    dx = MaxPooling2DFn.backward(c, T(dy))
    pool.update(dict(nan_x=x, nan_y=npy(y), nan_dy=dy, nan_dx=npy(dx)))
    np.savez_compressed(os.path.join(OUT, "pool.npz"), **pool)

    # ---------------- batchnorm 1-D/2-D (normalizations_test.py:11: eps {1e-5,1e-4} x m {0.1,0.2})
    bn = {}
    bcases = [((4, 3, 5, 5), 1e-5, 0.1, True), ((4, 3, 5, 5), 1e-4, 0.2, True), ((2, 6, 4, 4), 1e-5, 0.1, False),
              ((6, 5), 1e-5, 0.1, True), ((4, 5, 7), 1e-5, 0.2, True), ((6, 5), 1e-5, 0.1, False)]
    for n, (shape, eps, m, training) in enumerate(bcases):
        C = shape[1]
        x = rnd(shape, -1.0, 3.0)  # non-zero mean: exercises the variance cancellation
        w = rnd((C,), 0.5, 1.5, seed=44)
        b = rnd((C,), -0.5, 0.5, seed=45)
        rmean = rnd((C,), -0.2, 0.2, seed=46)
        rvar = rnd((C,), 0.5, 1.5, seed=47)
        Fn = BatchNorm2DFn if len(shape) == 4 else BatchNorm1DFn
        c = FunctionCache()
        y, rm2, rv2 = Fn.forward(c, T(x), T(rmean), T(rvar), T(w), T(b), m, eps, training)
        dy = rnd(shape, seed=43)
        dx, dw, db = Fn.backward(c, T(dy))
        for k, v in dict(x=x, w=w, b=b, rmean=rmean, rvar=rvar, y=npy(y), rmean2=npy(rm2), rvar2=npy(rv2),
                         dy=dy, dx=npy(dx), dw=npy(dw), db=npy(db)).items():
            bn[f"c{n}_{k}"] = v
        manifest.setdefault("batchnorm", []).append(dict(id=n, shape=list(shape), eps=eps, m=m, training=training))
    np.savez_compressed(os.path.join(OUT, "batchnorm.npz"), **bn)

    # ---------------- relu / dropout / cross-entropy
    misc = {}
    x = rnd((3, 4, 5, 5)); x[0, 0, 0, :3] = 0.0
    c = FunctionCache(); y = ReLUFn.forward(c, T(x)); dy = rnd(x.shape, seed=43); dx = ReLUFn.backward(c, T(dy))
    misc.update(relu_x=x, relu_y=npy(y), relu_dy=dy, relu_dx=npy(dx))
    x = rnd((4, 6, 3, 3))
    cp.random.set_seed(7)
    c = FunctionCache(); y = DropoutFn.forward(c, T(x), 0.25, True); dx = DropoutFn.backward(c, T(dy := rnd(x.shape, seed=43)))
    cp.random.set_seed(None)
    misc.update(drop_x=x, drop_y=npy(y), drop_dy=dy, drop_dx=npy(dx))
    for n, (B, NC) in enumerate([(6, 10), (5, 37)]):
        logits = rnd((B, NC), -3, 3)
        np.random.seed(42); targets = np.random.randint(0, NC, (B,)).astype(np.int64)
        c = FunctionCache()
        loss = CrossEntropyLossFn.forward(c, T(logits), T(targets), 1e-8)
        dl = CrossEntropyLossFn.backward(c)
        misc.update({f"ce{n}_logits": logits, f"ce{n}_targets": targets, f"ce{n}_loss": np.asarray(loss.item(), np.float32),
                     f"ce{n}_dlogits": npy(dl)})
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **misc)

    # ---------------- optimizers: 5 steps (tests/nn/test_optimizers.py:18-168)
    opt = {}
    ocases = [("sgd", dict(lr=0.1)), ("sgd", dict(lr=0.1, momentum=0.9)),
              ("sgd", dict(lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.1)),
              ("adam", dict(lr=1e-2)), ("adam", dict(lr=1e-2, weight_decay=0.1, beta1=0.8, beta2=0.99)),
              ("adamw", dict(lr=1e-2, weight_decay=0.1)), ("nadam", dict(lr=1e-2)),
              ("nadam", dict(lr=1e-2, weight_decay=0.05, momentum_decay=1e-2))]
    for n, (name, kw) in enumerate(ocases):
        p0 = [rnd((4, 6), -1, 1), rnd((5,), -1, 1, seed=48)]
        params = [Parameter(T(a.copy())) for a in p0]
        O = {"sgd": cp.nn.optimizers.SGD, "adam": cp.nn.optimizers.Adam, "adamw": cp.nn.optimizers.AdamW,
             "nadam": cp.nn.optimizers.NAdam}[name]
        o = O(params, **kw)
        for step in range(5):
            for j, p in enumerate(params):
                g = rnd(p.shape, -1, 1, seed=100 + 10 * step + j)
                opt[f"c{n}_g{step}_{j}"] = g
                p.grad = T(g.copy())
            o.step()
            for j, p in enumerate(params):
                opt[f"c{n}_p{step}_{j}"] = p.to_numpy().copy()
        for j, a in enumerate(p0):
            opt[f"c{n}_init_{j}"] = a
        manifest.setdefault("optim", []).append(dict(id=n, name=name, kw=kw))
    np.savez_compressed(os.path.join(OUT, "optim.npz"), **opt)

    # ---------------- a 2-step train trace of a tiny CNN through the reference's modules (model-level pin)
    from compyute.nn import (BatchNorm2D, Conv2D, CrossEntropyLoss, Flatten, Linear, MaxPooling2D, ReLU, Sequential)
    from compyute.nn.optimizers import Adam
    cp.random.set_seed(3)
    # (no conv bias in front of BatchNorm: its gradient is exactly-zero-plus-rounding-noise, which Adam amplifies to +-lr)
    model = Sequential(Conv2D(2, 4, 3, padding="same", bias=False), BatchNorm2D(4), ReLU(), MaxPooling2D(2),
                       Conv2D(4, 6, 3, padding="valid"), ReLU(), Flatten(), Linear(6 * 2 * 2, 5))
    cp.random.set_seed(None)
    model.training()
    tr = {f"init_{k}": v.to_numpy().copy() for k, v in model.get_state_dict().items()}
    x = rnd((6, 2, 8, 8), -1, 1)
    np.random.seed(42); t = np.random.randint(0, 5, (6,)).astype(np.int64)
    tr.update(x=x, t=t)
    lossf = CrossEntropyLoss(); optim = Adam(model.get_parameters(), lr=1e-2)
    for step in range(2):
        y = model(T(x)); loss = lossf(y, T(t)).item()
        optim.reset_grads(); model.backward(lossf.backward()); optim.step()
        tr[f"loss{step}"] = np.asarray(loss, np.float32); tr[f"logits{step}"] = y.to_numpy().copy()
    tr.update({f"final_{k}": v.to_numpy().copy() for k, v in model.get_state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "train_trace.npz"), **tr)
    manifest["train_trace"] = [dict(keys=list(model.get_state_dict().keys()))]

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    total = sum(os.path.getsize(os.path.join(OUT, n)) for n in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} files, {total/1024:.1f} KiB, reference v{cp.__version__}")


if __name__ == "__main__":
    main()
