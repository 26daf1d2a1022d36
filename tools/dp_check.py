"""Data-parallel parity check, launched with torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Every rank trains a small CNN (no BatchNorm, so DP with equal shards is mathematically the single-GPU full-batch run)
for 3 Adam steps on its shard; rank 0 also runs the full batch alone with gradient sync disabled and compares the
parameters (1e-5) and the loss trace.  Prints DP_CHECK OK / FAIL.

DP_SYNCBN=1: the model contains BatchNorm2D (+ fused ReLU) and BatchNorm1D layers and synchronised BatchNorm is switched on
(``distributed.set_sync_batchnorm``): statistics and backward sums over the global batch make DP on shards equal to the
single-GPU full-batch run again (SGD with momentum, 1e-4; per-shard statistics would differ at the 1e-1 level).  DP_MODE
selects the compute mode (fp32 default; bf16 exercises the channels-last producer paths).  DP_CLIP=<max_norm>: gradient clipping
before every step — the data-parallel run must clip the averaged global gradient to stay equal to the single-process run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compyute_b200 as cp
from compyute_b200 import distributed as D
from compyute_b200 import nn


SYNCBN = os.environ.get("DP_SYNCBN", "0") == "1"
MODE = os.environ.get("DP_MODE", "fp32")
CLIP = float(os.environ.get("DP_CLIP", "0"))  # > 0: clip_grad_norm(max_norm) before every step, in the DP and the single-process run


def build():
    np.random.seed(7)
    with cp.use_device(cp.cuda):
        if SYNCBN:
            return nn.Sequential(nn.Conv2D(3, 8, 3, padding="same"), nn.BatchNorm2D(8), nn.ReLU(), nn.MaxPooling2D(2),
                                 nn.Conv2D(8, 16, 3, padding="same", bias=False), nn.BatchNorm2D(16), nn.AvgPooling2D(2), nn.Flatten(),
                                 nn.Linear(16 * 4 * 4, 32, bias=False), nn.BatchNorm1D(32), nn.ReLU(), nn.Linear(32, 10))
        return nn.Sequential(nn.Conv2D(3, 8, 3, padding="same"), nn.ReLU(), nn.MaxPooling2D(2),
                             nn.Conv2D(8, 16, 3, padding="same", bias=False), nn.ReLU(), nn.AvgPooling2D(2), nn.Flatten(),
                             nn.Linear(16 * 4 * 4, 32), nn.ReLU(), nn.Linear(32, 10))


def train(model, x, t, steps, dp):
    model.training()
    opt = nn.optimizers.SGD(model.get_parameters(), lr=5e-2, momentum=0.9) if SYNCBN else nn.optimizers.Adam(model.get_parameters(), lr=1e-2)
    opt._data_parallel = dp
    opt.fused_dp_step = dp and os.environ.get("DP_FUSED", "1") == "1"  # NVLS / peer-memory fused exchange + update (default)
    opt.overlap_grad_sync = dp and os.environ.get("DP_OVERLAP", "0") == "1"  # bucketed all-reduces during backward
    opt.bucket_bytes = int(os.environ.get("DP_BUCKET_BYTES", 4096))           # tiny buckets: several of them even in this small model
    loss_fn = nn.CrossEntropyLoss()
    xt, tt = cp.tensor(x, device=cp.cuda), cp.tensor(t, device=cp.cuda)
    losses = []
    D.set_sync_batchnorm(SYNCBN and dp)
    with cp.compute_mode(MODE):
        steps = max(steps, int(os.environ.get("DP_STEPS", steps)))
    for _ in range(steps):
            loss = loss_fn(model(xt), tt)
            opt.reset_grads()
            model.backward(loss_fn.backward())
            if CLIP > 0.0:  # clipping must see the GLOBAL-batch gradient on every rank (Optimizer.sync_grads)
                nn.utils.clip_grad_norm(model.get_parameters(), CLIP)
            opt.step()
            losses.append(loss.item())
    D.set_sync_batchnorm(False)
    if dp and D.rank() == 0:
        print("gradient exchange:", "fused step, " + opt.fused_dp_note() if opt._fused is not None else "NCCL all-reduce + replicated update (" + opt.fused_dp_note() + ")")
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
            tol = (1e-4 if MODE == "fp32" else 2e-2) if SYNCBN else 1e-5
            good = np.allclose(a.to_numpy(), b.to_numpy(), rtol=tol, atol=tol)
            ok &= bool(good)
            print(f"{k:14s} max|dp - single| = {err:.2e} {'ok' if good else 'MISMATCH'}")
        ok &= bool(np.allclose(lt.cpu().numpy(), ref_losses, rtol=tol, atol=tol))
        print("losses dp", lt.cpu().numpy().tolist(), "single", ref_losses)
    if SYNCBN:  # function-level check on UNEVEN shards: rank r holds 3 + 2*(r % 2) samples; y, running stats and dx must equal the
        # corresponding slice of the single-process evaluation on the concatenated batch (1e-5)
        from compyute_b200.nn.functional import BatchNorm2DFn, BatchNormReLU2DFn, FunctionCache
        sizes = [3 + 2 * (r % 2) for r in range(world)]
        lo = sum(sizes[:rank]); n = sizes[rank]
        xa = rng.normal(0.5, 2, (sum(sizes), 6, 5, 5)).astype(np.float32); dya = rng.normal(0, 1, xa.shape).astype(np.float32)
        wv = rng.uniform(0.5, 1.5, (6,)).astype(np.float32); bv = rng.uniform(-0.5, 0.5, (6,)).astype(np.float32)
        T = lambda a: cp.tensor(a, device=cp.cuda)
        for Fn in (BatchNorm2DFn, BatchNormReLU2DFn):
            res = {}
            for tag, sync, xs_, dys_ in (("full", False, xa, dya), ("shard", True, xa[lo:lo + n], dya[lo:lo + n])):
                D.set_sync_batchnorm(sync)
                c = FunctionCache()
                y, rm, rv = Fn.forward(c, T(xs_), T(np.zeros(6, np.float32)), T(np.ones(6, np.float32)), T(wv), T(bv), 0.1, 1e-5, True)
                dx, dw, db = Fn.backward(c, T(dys_))
                res[tag] = [a.to_numpy() for a in (y, dx, rm, rv)]
                D.set_sync_batchnorm(False)
            for name, full, shard in zip(("y", "dx", "rmean", "rvar"), res["full"], res["shard"]):
                ref = full[lo:lo + n] if name in ("y", "dx") else full
                good = bool(np.allclose(shard, ref, rtol=1e-5, atol=1e-5))
                ok &= good
                if rank == 0 or not good:
                    print(f"syncbn {Fn.__name__:18s} {name:5s} rank {rank} max|shard - full| = {np.abs(shard - ref).max():.2e} {'ok' if good else 'MISMATCH'}")
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
