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


# ---------------------------------------------------------------- synchronised BatchNorm: the exchange protocol, on CPU
def _syncbn_worker(rank, world, port, q):
    """What the SyncBN path does around its two collectives (normalization_funcs._bn_forward/_bn_backward + bn.cu), restated
    in NumPy on each rank's UNEVEN shard and exchanged over gloo: local (mean, M2, n) -> all-gather -> Chan merge in rank order;
    local (Σdy, Σdy·x̂) -> SUM all-reduce -> dx with the global count.  Must reproduce the oracle's full-batch BatchNorm."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from compyute_b200 import distributed as D
    D.init("gloo")
    D.set_sync_batchnorm(True)
    active = D.sync_batchnorm_active()
    D.set_sync_batchnorm(False)
    ok = active and not D.sync_batchnorm_active()
    rng = np.random.RandomState(1)
    sizes = [5, 3]
    x = rng.normal(0.5, 2.0, (sum(sizes), 6, 4, 4)).astype(np.float32); dy = rng.normal(0, 1, x.shape).astype(np.float32)
    w = rng.uniform(0.5, 1.5, 6).astype(np.float32); b = rng.uniform(-0.5, 0.5, 6).astype(np.float32)
    lo = sum(sizes[:rank]); xs, dys = x[lo:lo + sizes[rank]], dy[lo:lo + sizes[rank]]
    # forward: cpt_bn_local_stats -> all_gather -> bn_fwd_finalize_merged_kernel
    n_loc = xs.shape[0] * 16
    mean_loc = xs.mean((0, 2, 3), dtype=np.float64); m2_loc = ((xs - mean_loc[None, :, None, None]) ** 2).sum((0, 2, 3), dtype=np.float64)
    stats = torch.from_numpy(np.stack([mean_loc, m2_loc, np.full(6, n_loc)]).astype(np.float32))
    gathered = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)
    n = 0.0; mean = np.zeros(6); m2 = np.zeros(6)
    for g in gathered:  # rank order, Chan's pairwise update
        g = g.numpy().astype(np.float64); nb = g[2, 0]
        delta = g[0] - mean; tot = n + nb
        mean = mean + delta * nb / tot; m2 = m2 + g[1] + delta ** 2 * n * nb / tot; n = tot
    rstd = 1.0 / np.sqrt(m2 / n + 1e-5)
    xhat = (xs - mean[None, :, None, None]) * rstd[None, :, None, None]
    y = w[None, :, None, None] * xhat + b[None, :, None, None]
    # backward: cpt_bn_act_bwd_local_sums -> all_reduce -> cpt_bn_act_bwd_apply_global
    sums = torch.from_numpy(np.stack([dys.sum((0, 2, 3), dtype=np.float64), (dys * xhat).sum((0, 2, 3), dtype=np.float64)]))
    dw_loc, db_loc = sums[1].numpy().copy(), sums[0].numpy().copy()
    dist.all_reduce(sums)
    s1, s2 = sums[0].numpy()[None, :, None, None], sums[1].numpy()[None, :, None, None]
    dx = (w * rstd)[None, :, None, None] * (dys - s1 / n - xhat * s2 / n)
    # oracle on the concatenated batch
    rc = []
    y_ref, rm_ref, rv_ref = R.batchnorm_forward(rc, x, np.zeros(6, np.float32), np.ones(6, np.float32), w, b, 0.1, 1e-5, True)
    dx_ref, dw_ref, db_ref = R.batchnorm_backward(rc, dy)
    ok = ok and np.allclose(y, y_ref[lo:lo + sizes[rank]], rtol=1e-5, atol=1e-5) and np.allclose(dx, dx_ref[lo:lo + sizes[rank]], rtol=1e-5, atol=1e-5)
    ok = ok and np.allclose(0.9 * 0 + 0.1 * mean, rm_ref, rtol=1e-5, atol=1e-6) and np.allclose(0.9 + 0.1 * m2 / (n - 1), rv_ref, rtol=1e-5, atol=1e-6)
    # dw / db stay local sums; their SUM over ranks is the full-batch gradient (the gradient exchange then applies its 1/world,
    # matching the 1/world of the averaged loss)
    g = torch.from_numpy(np.stack([dw_loc, db_loc])); dist.all_reduce(g)
    ok = ok and np.allclose(g[0].numpy(), dw_ref.ravel(), rtol=1e-5, atol=1e-5) and np.allclose(g[1].numpy(), db_ref.ravel(), rtol=1e-5, atol=1e-5)
    D.barrier()
    q.put((rank, bool(ok)))


def test_syncbn_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_plan_shards_for_the_fused_step():
    """Layout of the symmetric gradient / parameter arenas of the fused data-parallel step (csrc/dp_step.cu): equal
    contiguous shards, 16-byte vectors never straddle a shard, padding only at the end, every element owned exactly once."""
    from compyute_b200.distributed import plan_shards
    for total in (1, 63, 64, 65, 1000, 12_345_677, 134_250_496):
        for world in (1, 2, 3, 4, 8):
            padded, shard = plan_shards(total, world)
            assert padded == shard * world and padded >= total and shard % 64 == 0
            assert padded - total < world * 64  # at most one alignment unit of padding per rank
            owners = [(r * shard, (r + 1) * shard) for r in range(world)]
            assert owners[0][0] == 0 and owners[-1][1] == padded and all(a[1] == b[0] for a, b in zip(owners, owners[1:]))


def test_shard_weight_restores_the_global_mean():
    """Uneven shards (shard_bounds spreads the remainder over the first ranks): averaging the ranks' local-mean gradients with
    weights n_r * world / N before the sum / world exchange gives the global-batch mean exactly."""
    from compyute_b200.distributed import shard_bounds, shard_weight
    rng = np.random.RandomState(0)
    g = rng.normal(0, 1, (11, 5))  # per-sample gradients
    for world in (2, 3, 4):
        acc = np.zeros(5)
        for r in range(world):
            lo, hi = shard_bounds(11, r, world)
            acc += shard_weight(hi - lo, 11, world) * g[lo:hi].mean(0)
        assert np.allclose(acc / world, g.mean(0), rtol=1e-12, atol=1e-12)
    assert shard_weight(8, 32, 4) == 1.0
