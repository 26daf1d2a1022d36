"""world_size-2 gloo tests (CPU): the data-parallel plumbing — sharding, the single SUM all-reduce of the gradient
arena and the 1/world scaling folded into the update — reproduces the reference's full-batch gradient.
The oracle (per-shard reference run, gradients averaged) is the checker; no CUDA is involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import compyute_ref as R


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _grads(x, t, w, b):
    """Linear + cross-entropy forward/backward through the oracle; returns (dw, db) for the local batch."""
    c = []
    logits = R.linear_forward(c, x, w, b)
    lc = []
    R.cross_entropy_forward(lc, logits, t)
    dl = R.cross_entropy_backward(lc)
    _, dw, db = R.linear_backward(c, dl)
    return dw, db


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from compyute_b200 import distributed as D
    D.init("gloo")
    assert D.is_initialized() and D.world_size() == world and D.rank() == rank
    rng = np.random.RandomState(0)
    x = rng.normal(0, 1, (12, 7)).astype(np.float32); t = rng.randint(0, 5, (12,))
    w = rng.uniform(-0.3, 0.3, (5, 7)).astype(np.float32); b = rng.uniform(-0.3, 0.3, (5,)).astype(np.float32)
    lo, hi = D.shard_bounds(12)
    xs, ts = D.shard_batch(x, t)
    assert xs.shape[0] == hi - lo == 6 and np.array_equal(xs, x[lo:hi])
    dw, db = _grads(xs, ts, w, b)
    arena = torch.from_numpy(np.concatenate([dw.ravel(), db.ravel()]).astype(np.float32))  # flat gradient arena
    D.all_reduce_sum(arena)                                                                # the path's one collective
    avg = arena.numpy() * (1.0 / world)                                                    # grad_scale of the fused step
    dw_full, db_full = _grads(x, t, w, b)
    ok = np.allclose(avg, np.concatenate([dw_full.ravel(), db_full.ravel()]), rtol=1e-5, atol=1e-6)
    # broadcast keeps replicas identical
    p = torch.full((4,), float(rank))
    class _T:  # minimal stand-in for a Tensor with a DeviceArray-like .data._buf
        class data: _buf = p
    D.broadcast_parameters([_T], src=0)
    ok = ok and bool((p == 0).all())
    D.barrier()
    q.put((rank, bool(ok), (lo, hi)))


def test_dp_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (0, 6)), (1, True, (6, 12))]


def test_shard_bounds_cover_batch():
    from compyute_b200.distributed import shard_bounds
    for n in (1, 7, 8, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
