"""Where does a model-level run in a tensor-core mode leave the rounded-operand oracle?  VGG-style model (width 32, 32x32,
B=4): forward activations after every layer (fusion off, walked layer by layer) and final logits / first gradients with
fusion and epilogue statistics on/off, against oracle/model_ref.RefModel(operand_rounding=mode).  Run on the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as W
import compyute_b200 as cp
from compyute_b200 import nn
from compyute_b200.nn.modules.containers import set_epilogue_stats_enabled, set_fusion_enabled
from oracle import compyute_ref as R
from oracle.model_ref import RefModel

spec = W.vgg(width=32, hw=32, hidden=64)
rng = np.random.RandomState(5)
x = rng.normal(0, 1, (4, 3, 32, 32)).astype(np.float32)
t = rng.randint(0, 10, (4,))
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

for mode in ("bf16", "tf32"):
    np.random.seed(11)
    with cp.use_device(cp.cuda):
        model = W.build(spec)
    model.training()
    params0 = [p.to_numpy().copy() for p in model.get_parameters()]
    bufs0 = [b.to_numpy().copy() for b in model.get_buffers()]
    ref = RefModel(spec, [p.copy() for p in params0], [b.copy() for b in bufs0], operand_rounding=mode)
    acts, a = [], x
    for l in ref.layers:
        a = l.forward(a, True)
        acts.append(a)
    lc = []
    loss_ref = float(R.cross_entropy_forward(lc, a, t))
    ref.backward(R.cross_entropy_backward(lc))
    g_ref = ref.gradients()
    xt, tt = cp.tensor(x, device=cp.cuda), cp.tensor(t.astype(np.int32), device=cp.cuda)
    with cp.compute_mode(mode):
        set_fusion_enabled(False)
        h = xt
        for i, (l, s) in enumerate(zip(model.layers, spec)):
            h = l(h)
            print(f"{mode} layer {i:2d} {s[0]:8s} fwd rel err vs rounded oracle {rel(h.to_numpy(), acts[i]):.3e}", flush=True)
        for l in model.get_modules():
            l.fcache.cache.clear()
        for fusion, stats in ((False, False), (True, False), (True, True)):
            set_fusion_enabled(fusion); set_epilogue_stats_enabled(stats)
            for p, p0 in zip(model.get_parameters(), params0):
                p.grad = None
            loss_fn = nn.CrossEntropyLoss()
            y = model(xt)
            loss = loss_fn(y, tt)
            model.backward(loss_fn.backward())
            gerr = [float(np.linalg.norm(p.grad.to_numpy() - g) / max(np.linalg.norm(g), 1e-30)) for p, g in zip(model.get_parameters(), g_ref)
                    if np.abs(g).max() > 1e-5 * max(np.abs(q).max() for q in g_ref)]
            print(f"{mode} fusion={fusion} epilogue_stats={stats}: logits rel {rel(y.to_numpy(), acts[-1]):.3e} loss {loss.item():.7f} vs {loss_ref:.7f}"
                  f" grad relL2 median {np.median(gerr):.3e} max {max(gerr):.3e}", flush=True)
        set_fusion_enabled(True); set_epilogue_stats_enabled(True)
