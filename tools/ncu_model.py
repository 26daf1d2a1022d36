"""Profiling driver for the model workloads: builds $NCU_WORKLOAD (vgg | resnet18 | mnist | mlp), runs $NCU_WARM warm-up
train steps, then ONE step between cudaProfilerStart/Stop.  Use with
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/ncu_model.py
so that only that step's launches are listed (a whole bench.py run under ncu costs minutes of GPU time)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as W  # noqa: E402
import compyute_b200 as cp  # noqa: E402
from bench import MODEL_WORKLOADS  # noqa: E402
from compyute_b200 import nn  # noqa: E402
from compyute_b200.tensors import DeviceArray, Tensor  # noqa: E402

name = os.environ.get("NCU_WORKLOAD", "resnet18")
factory, xshape, classes, B, _ = MODEL_WORKLOADS[name]
B = int(os.environ.get("NCU_B", B))
mode = os.environ.get("NCU_MODE", "bf16")
np.random.seed(0)
with cp.use_device(cp.cuda), cp.compute_mode(mode):
    model = W.build(getattr(W, factory)())
model.training()
opt = nn.optimizers.Adam(model.get_parameters(), lr=1e-3)
loss_fn = nn.CrossEntropyLoss()
x = torch.randn(B, *xshape, device="cuda")
t = torch.randint(0, classes, (B,), dtype=torch.int32, device="cuda")
wx = Tensor(DeviceArray(x, tuple(x.shape), np.float32))
wt = Tensor(DeviceArray(t, tuple(t.shape), np.int32))


def step():
    loss = loss_fn(model(wx), wt)
    opt.reset_grads()
    model.backward(loss_fn.backward())
    opt.step()
    return loss


with cp.compute_mode(mode):
    for _ in range(int(os.environ.get("NCU_WARM", 2))):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    loss = step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("loss", loss.item())
