"""Golden vectors for the widened ("next") rows of SURVEY §8f, from the REAL reference (TEST INFRASTRUCTURE ONLY; run in the
build container, like oracle/gen_golden.py whose import shims it reuses):

    python oracle/gen_golden_next.py       # writes tests/golden/lr_schedulers.json, tests/golden/clip_grad_norm.npz

* learning-rate sequences of the five schedulers of compyute/nn/utils/lr_schedulers.py driven by a bare optimizer whose step
  counter advances like ``Optimizer.step`` does (optimizers.py:29, 176: t starts at 1, +1 per step);
* ``clip_grad_norm`` (compyute/nn/utils/training.py:12-39) on seeded gradients, clipped and unclipped.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.gen_golden import OUT, _install_shims  # noqa: E402

LR_CASES = [
    ("StepLrScheduler", {"t_decay": 7, "lr_decay": 0.1}),
    ("MultistepLrScheduler", {"t_decay_step": 5, "lr_decay": 0.5}),
    ("ExponentialLrScheduler", {"decay_steps": 12, "lr_decay": 0.9}),
    ("CosineLrScheduler", {"target_lr": 1e-4, "warmup_steps": 6, "decay_steps": 15}),
    ("AdaptiveLrScheduler", {"patience": 4, "lr_downscale_factor": 0.5, "lr_upscale_factor": 1.5}),
]
STEPS = 30


def metric_series(n: int) -> list[float]:
    rng = np.random.RandomState(3)
    return [float(v) for v in (np.linspace(2.0, 1.0, n) + rng.normal(0, 0.15, n))]


def main() -> None:
    _install_shims()
    import compyute as cp
    from compyute.nn import optimizers
    from compyute.nn.parameter import Parameter
    from compyute.nn.utils import lr_schedulers
    from compyute.nn.utils.training import clip_grad_norm

    out = []
    for name, kw in LR_CASES:
        opt = optimizers.SGD(lr=0.01)
        opt.parameters = []
        sched = getattr(lr_schedulers, name)(opt, **kw)
        ms = metric_series(STEPS)
        for i in range(STEPS):
            opt.t += 1  # what Optimizer.step does after the update
            if name == "AdaptiveLrScheduler":
                sched.step(loss=ms[i])
            else:
                sched.step()
        out.append({"scheduler": name, "kwargs": kw, "lr0": 0.01, "steps": STEPS, "lr_history": sched.cache["lr_history"],
                    "final_lr": opt.lr, "metrics": ms if name == "AdaptiveLrScheduler" else None})
    with open(os.path.join(OUT, "lr_schedulers.json"), "w") as f:
        json.dump(out, f, indent=1)

    rng = np.random.RandomState(42)
    shapes = [(8, 3, 3, 3), (8,), (16, 72), (16,), (5, 16)]
    grads = [rng.uniform(-0.1, 0.1, s).astype(np.float32) for s in shapes]
    res = {f"g{i}": g for i, g in enumerate(grads)}
    for tag, max_norm in (("loose", 10.0), ("tight", 0.25)):
        ps = []
        for g in grads:
            p = Parameter(cp.tensor(np.zeros_like(g)))
            p.grad = cp.tensor(g.copy())
            ps.append(p)
        res[f"{tag}_max_norm"] = np.float64(max_norm)
        res[f"{tag}_norm"] = np.float64(clip_grad_norm(iter(ps), max_norm))
        for i, p in enumerate(ps):
            res[f"{tag}_g{i}"] = p.grad.to_numpy()
    np.savez(os.path.join(OUT, "clip_grad_norm.npz"), **res)
    print("wrote lr_schedulers.json, clip_grad_norm.npz")


if __name__ == "__main__":
    main()
