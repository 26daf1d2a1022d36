"""Data-parallel parity check, launched with torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Every rank trains a small CNN (no BatchNorm, so DP with equal shards is mathematically the single-GPU full-batch run)
for 3 Adam steps on its shard; rank 0 also runs the full batch alone with gradient sync disabled and compares the
parameters (1e-5) and the loss trace.  Prints DP_CHECK OK / FAIL."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200 import distributed as D
from compyute_b200 import nn


def build():
    np.random.seed(7)
    with cp.use_device(cp.cuda):
        return nn.Sequential(nn.Conv2D(3, 8, 3, padding="same"), nn.ReLU(), nn.MaxPooling2D(2),
                             nn.Conv2D(8, 16, 3, padding="same", bias=False), nn.ReLU(), nn.AvgPooling2D(2), nn.Flatten(),
                             nn.Linear(16 * 4 * 4, 32), nn.ReLU(), nn.Linear(32, 10))


def train(model, x, t, steps, dp):
    model.training()
    opt = nn.optimizers.Adam(model.get_parameters(), lr=1e-2)
    opt._data_parallel = dp
    opt.overlap_grad_sync = dp and os.environ.get("DP_OVERLAP", "0") == "1"  # bucketed all-reduces during backward
    opt.bucket_bytes = int(os.environ.get("DP_BUCKET_BYTES", 4096))           # tiny buckets: several of them even in this small model
    loss_fn = nn.CrossEntropyLoss()
    xt, tt = cp.tensor(x, device=cp.cuda), cp.tensor(t, device=cp.cuda)
    losses = []
    for _ in range(steps):
        loss = loss_fn(model(xt), tt)
        opt.reset_grads()
        model.backward(loss_fn.backward())
        opt.step()
        losses.append(loss.item())
    return losses


if __name__ == "__main__":
    D.init("nccl")
    rank, world = D.rank(), D.world_size()
    rng = np.random.RandomState(0)
    B = 8 * world
    x = rng.normal(0, 1, (B, 3, 16, 16)).astype(np.float32)
    t = rng.randint(0, 10, (B,)).astype(np.int32)
    model = build()
    D.broadcast_parameters(model.get_state_dict().values())
    xs, ts = D.shard_batch(x, t)
    losses = train(model, xs, ts, 3, dp=True)
    # global loss = mean of the shard losses
    lt = torch.tensor(losses, device="cuda"); torch.distributed.all_reduce(lt); lt /= world
    ok = True
    if rank == 0:
        ref = build()
        ref_losses = train(ref, x, t, 3, dp=False)
        for (k, a), (_, b) in zip(model.get_state_dict().items(), ref.get_state_dict().items()):
            err = np.abs(a.to_numpy() - b.to_numpy()).max()
            good = np.allclose(a.to_numpy(), b.to_numpy(), rtol=1e-5, atol=1e-5)
            ok &= bool(good)
            print(f"{k:14s} max|dp - single| = {err:.2e} {'ok' if good else 'MISMATCH'}")
        ok &= bool(np.allclose(lt.cpu().numpy(), ref_losses, rtol=1e-5, atol=1e-6))
        print("losses dp", lt.cpu().numpy().tolist(), "single", ref_losses)
    # every replica holds identical parameters after the steps
    for k, v in model.get_state_dict().items():
        a = v.data._buf.clone(); b = a.clone()
        torch.distributed.broadcast(b, src=0)
        ok &= bool(torch.equal(a, b))
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda"); torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print("DP_CHECK", "OK" if flag.item() == 1.0 else "FAIL", f"world={world}")
    D.barrier()
    sys.exit(0 if flag.item() == 1.0 else 1)
