"""Times the data-parallel exchange + update alone on N GPUs (torchrun): the fused sharded step (barrier, kernel, barrier) for a
537 MB gradient arena (the MLP config: 134.3 M parameters) against the classic NCCL all-reduce + replicated Adam update.
Launch shape comes from the environment (CPT_DP_UNROLL, CPT_DP_CTAS_PER_SM, CPT_DP_MULTIMEM): one process group per setting.
    python -m torch.distributed.run --nproc-per-node 8 ... tools/dp_step_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200 import distributed as D
from compyute_b200 import nn
from compyute_b200.tensors import DeviceArray, Tensor

D.init("nccl")
rank, world = D.rank(), D.world_size()
n_layers, width = 8, 4096
params = []
with cp.use_device(cp.cuda):
    for _ in range(n_layers):
        params.append(nn.Parameter(cp.tensor(np.zeros((width, width), np.float32), device=cp.cuda)))
        params.append(nn.Parameter(cp.tensor(np.zeros((width,), np.float32), device=cp.cuda)))


def run(fused: bool, steps=12, warm=3):
    opt = nn.optimizers.Adam(params, lr=1e-3)
    opt.fused_dp_step = fused
    for p in params:
        p.grad = Tensor(p.grad_slot)
        p.grad_slot._buf.normal_()
    ts = []
    for i in range(steps + warm):
        for p in params:
            if p.grad is None or p.grad.data.ptr != p.grad_slot.ptr:
                p.grad = Tensor(p.grad_slot)
        torch.cuda.synchronize(); D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); opt.step(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        if i >= warm:
            ts.append(float(t.item()))
    return float(np.median(ts)), float(np.min(ts)), opt.fused_dp_note()


res = {"world": world, "bytes": sum(p.size for p in params) * 4, "unroll": os.environ.get("CPT_DP_UNROLL", "4"),
       "ctas_per_sm": os.environ.get("CPT_DP_CTAS_PER_SM", "4"), "multimem_env": os.environ.get("CPT_DP_MULTIMEM")}
med, mn, note = run(True)
res.update(fused_ms_median=round(med, 3), fused_ms_min=round(mn, 3), fused=note)
if os.environ.get("DP_BENCH_NCCL", "0") == "1":
    med, mn, _ = run(False)
    res.update(nccl_plus_adam_ms_median=round(med, 3), nccl_plus_adam_ms_min=round(mn, 3))
if rank == 0:
    print(json.dumps(res), flush=True)
D.barrier()
