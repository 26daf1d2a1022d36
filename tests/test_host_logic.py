"""CPU-only tests: the C ABI exports what include/*.h declares, host-side protocol logic (cache, module registry,
state dict, optimizer bookkeeping), and the no-fallback guarantees.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN

import compyute_b200 as cp
from compyute_b200 import _lib, nn
from compyute_b200.nn.functional import FunctionCache, PseudoCache, conv2d, no_caching, relu


def test_abi_symbols_match_header():
    protos = _lib.parse_header()
    assert len(protos) >= 40
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/compyute_b200.h but not exported"
    # every extern "C" cpt_* symbol of the library is declared in the header (no undocumented entry points)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (cpt_\w+)", out))
    assert exported == set(protos), exported ^ set(protos)
    lib = _lib.lib()
    assert lib.cpt_version() == 100 and lib.cpt_last_error() is not None


def test_abi_argument_validation_without_gpu():
    """Argument errors are reported before any CUDA work, so they are testable on a CPU box."""
    lib = _lib.lib()
    d = _lib.ConvDesc(2, 3, 8, 8, 4, 3, 1, 1, 1)
    ho, wo = ctypes.c_int(), ctypes.c_int()
    assert lib.cpt_conv2d_out_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == 0 and (ho.value, wo.value) == (8, 8)
    d2 = _lib.ConvDesc(2, 3, 8, 8, 4, 3, 0, 2, 2)
    lib.cpt_conv2d_out_shape(ctypes.byref(d2), ctypes.byref(ho), ctypes.byref(wo))
    assert (ho.value, wo.value) == (2, 2)  # (8 - 2*2 - 1)//2 + 1
    bad = _lib.ConvDesc(2, 3, 2, 2, 4, 5, 0, 1, 1)
    assert lib.cpt_conv2d_out_shape(ctypes.byref(bad), ctypes.byref(ho), ctypes.byref(wo)) == _lib.ERR_INVALID
    assert b"larger than padded input" in lib.cpt_last_error()
    assert lib.cpt_maxpool2d_fwd(None, None, 1, 1, 2, 2, 3, None) == _lib.ERR_INVALID
    assert lib.cpt_conv2d_workspace_size(_lib.OP_WGRAD, ctypes.byref(d), _lib.MODE_FP32) > 0
    assert lib.cpt_channels_last_bytes(2, 3, 8, 8, _lib.MODE_BF16) >= 2 * 8 * 8 * 8 * 2
    with pytest.raises(cp.ShapeError):
        _lib.check(_lib.ERR_INVALID)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED)


def test_no_cpu_fallback():
    x = cp.tensor(np.zeros((2, 3, 8, 8), np.float32))
    w = cp.tensor(np.zeros((4, 3, 3, 3), np.float32))
    assert x.device == cp.cpu
    with pytest.raises(cp.DeviceError):
        conv2d(x, w)
    with pytest.raises(cp.DeviceError):
        relu(x)
    # nothing in the product package imports the oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "compyute_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, re.M), f"{f} imports the oracle"


def test_function_cache_protocol():
    c = FunctionCache()
    c.push(1, 2); c.push("a")
    assert c.pop() == ("a",) and c.pop() == (1, 2)
    with no_caching():
        c.push(3)
    assert not c.cache
    p = PseudoCache(); p.push(1)
    assert not p.cache


def test_module_registry_state_dict_and_modes():
    np.random.seed(0)
    m = nn.Sequential(nn.Conv2D(1, 2, 3, padding="same"), nn.BatchNorm2D(2), nn.ReLU(), nn.MaxPooling2D(2), nn.Flatten(),
                      nn.Linear(8, 3, bias=False))
    keys = list(m.get_state_dict().keys())
    assert keys == ["layers.0.w", "layers.0.b", "layers.1.w", "layers.1.b", "layers.1.rmean", "layers.1.rvar", "layers.5.w"]
    assert len(list(m.get_parameters())) == 5 and len(list(m.get_buffers())) == 2 and m.n_modules == 6
    conv = m.layers[0]
    assert conv.padding == 1 and conv.w.shape == (2, 1, 3, 3) and abs(conv.w.to_numpy()).max() <= 1 / 3
    assert nn.Conv2D(1, 1, 4, padding="same", dilation=2).padding == (4 * 2 - 1) // 2
    # default init is the reference's: first draw of U(-k, k) from numpy's legacy stream
    np.random.seed(0); w0 = np.random.uniform(-1 / 3, 1 / 3, (2, 1, 3, 3)).astype(np.float32)
    assert np.array_equal(conv.w.to_numpy(), w0)
    bn = m.layers[1]
    assert np.array_equal(bn.w.to_numpy(), np.ones(2)) and np.array_equal(bn.rvar.to_numpy(), np.ones(2))
    m.inference()
    assert all(isinstance(x.fcache, PseudoCache) and not x.is_training for x in m.get_modules())
    with pytest.raises(AttributeError):
        m.backward(cp.tensor(np.zeros((1, 3), np.float32)))
    m.training()
    assert all(type(x.fcache) is FunctionCache for x in m.get_modules())
    # load_state_dict: ordered, key-checked, rebinding .data
    m2 = nn.Sequential(nn.Conv2D(1, 2, 3, padding="same"), nn.BatchNorm2D(2), nn.ReLU(), nn.MaxPooling2D(2), nn.Flatten(),
                       nn.Linear(8, 3, bias=False))
    m2.load_state_dict(m.get_state_dict())
    assert m2.layers[0].w.data is m.layers[0].w.data
    with pytest.raises(ValueError):
        nn.Sequential(nn.BatchNorm2D(2)).load_state_dict(m.get_state_dict())
    with pytest.raises(nn.EmptyContainerError):
        nn.Sequential()
    m.trainable = False
    p = m.layers[5].w
    m.layers[5].update_parameter_grad(p, cp.tensor(np.ones(p.shape, np.float32)))
    assert p.grad is None
    m.trainable = True
    g = cp.tensor(np.ones(p.shape, np.float32))
    m.layers[5].update_parameter_grad(p, g)
    assert p.grad is g  # first grad stored by reference
    m.layers[5].update_parameter_grad(p, cp.tensor(np.ones(p.shape, np.float32)))
    assert p.grad is g and np.array_equal(g.to_numpy(), 2 * np.ones(p.shape))  # later ones accumulate in place
    m.clean()
    assert p.grad is None
    with pytest.raises(TypeError):
        nn.Parameter(cp.tensor(np.zeros(3, np.int32)))


def test_optimizer_bookkeeping():
    p = nn.Parameter(cp.tensor(np.ones((2, 2), np.float32)))
    q = nn.Parameter(cp.tensor(np.ones(3, np.float32)))
    o = nn.optimizers.Adam([p, q, p], lr=0.5)
    assert len(o._parameters) == 2 and o.t == 1 and o._state == {0: {}, 1: {}}
    sd = o.get_state_dict()
    assert set(sd) == {"state", "vars"} and sd["vars"]["lr"] == 0.5 and sd["vars"]["beta1"] == 0.9 and "t" in sd["vars"]
    o2 = nn.optimizers.Adam([p, q]); o2.load_state_dict({"state": {0: {}, 1: {}}, "vars": {"lr": 0.25, "t": 7}})
    assert o2.lr == 0.25 and o2.t == 7
    p.grad = cp.tensor(np.ones((2, 2), np.float32)); o.reset_grads()
    assert p.grad is None
    p.grad = cp.tensor(np.ones((2, 2), np.float32))
    with pytest.raises(TypeError):  # host parameters: no CPU fallback
        o.step()
    assert nn.optimizers.AdamW([p]).weight_decay == 1e-2 and nn.optimizers.SGD([p]).momentum == 0.0


def test_compute_mode_context():
    # "fp32" (the default) is the fp32-exact tensor-core mode; "fp32_simt" the FFMA kernels; both keep the 1e-5 contract
    assert cp.get_compute_mode() == _lib.MODE_FP32X3
    with cp.compute_mode("bf16"):
        assert cp.get_compute_mode() == _lib.MODE_BF16
        with cp.compute_mode("tf32"):
            assert cp.get_compute_mode() == _lib.MODE_TF32
        with cp.compute_mode("fp32_simt"):
            assert cp.get_compute_mode() == _lib.MODE_FP32
    assert cp.get_compute_mode() == _lib.MODE_FP32X3
    with cp.compute_mode("fp32x3"):
        assert cp.get_compute_mode() == _lib.MODE_FP32X3
    with cp.use_device(cp.cuda):
        assert cp.select_device(None) == cp.cuda
    assert cp.select_device(None) == cp.cpu


def test_dataloader_host_side():
    """Reference semantics of Dataloader (dataloaders.py:18-69): length, last partial batch, shuffling from numpy's
    legacy stream, and contiguous DP shards."""
    from compyute_b200.nn.utils import Dataloader
    x = cp.tensor(np.arange(10 * 3, dtype=np.float32).reshape(10, 3)); y = cp.tensor(np.arange(10, dtype=np.int64))
    dl = Dataloader((x, y), batch_size=4, shuffle_data=False)
    assert len(dl) == 3
    got = list(dl())
    assert [b[0].shape[0] for b in got] == [4, 4, 2] and got[2][1].to_numpy().tolist() == [8, 9]
    assert len(Dataloader((x, y), batch_size=4, drop_remaining=True)) == 2 and len(Dataloader((x, y), batch_size=64)) == 1
    np.random.seed(3); order = np.random.permutation(10)
    np.random.seed(3); first = next(iter(Dataloader((x, y), batch_size=5)()))
    assert first[1].to_numpy().tolist() == order[:5].tolist()
    assert got[0][1].dtype == np.int32  # labels are narrowed for the device kernels


def test_lr_scheduler_interplay_and_state_dict_roundtrip(tmp_path):
    """`lr` / `t` stay live python attributes (Appendix A.17) and optimizer/module state pickles through cp.save/load."""
    p = nn.Parameter(cp.tensor(np.ones((2, 2), np.float32)))
    o = nn.optimizers.NAdam([p], lr=0.1)
    for epoch in range(3):  # what compyute.nn.utils.lr_schedulers.ExponentialLrScheduler does: optimizer.lr *= decay
        o.lr *= 0.5
    assert abs(o.lr - 0.0125) < 1e-12
    sd = o.get_state_dict()
    assert sd["vars"]["lr"] == o.lr and "_mu_prod" in sd["vars"] and "momentum_decay" in sd["vars"]
    m = nn.Sequential(nn.Linear(3, 2), nn.ReLU())
    f = tmp_path / "state.cp"
    cp.save({"model": m.get_state_dict(), "optim": sd}, str(f))
    back = cp.load(str(f))
    assert list(back["model"].keys()) == ["layers.0.w", "layers.0.b"]
    assert np.array_equal(back["model"]["layers.0.w"].to_numpy(), m.layers[0].w.to_numpy())
    m2 = nn.Sequential(nn.Linear(3, 2), nn.ReLU()); m2.load_state_dict(back["model"])
    assert np.array_equal(m2.layers[0].w.to_numpy(), m.layers[0].w.to_numpy())


def test_plan_grad_buckets_tiles_the_arena_from_the_end():
    """Buckets of the overlapped data-parallel exchange: contiguous, disjoint, cover the arena, ordered from the last
    parameter (whose gradient backward produces first) to the first, each >= the requested size except the front one."""
    from compyute_b200.nn.optimizers import plan_grad_buckets
    rng = np.random.RandomState(0)
    for _ in range(20):
        sizes = [int(v) for v in rng.randint(1, 5000, rng.randint(1, 12))]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += (s + 63) // 64 * 64
        for be in (1, 700, 4096, 10 ** 7):
            buckets = plan_grad_buckets(offs, sizes, be)
            assert buckets[0][1] == tot and buckets[-1][0] == 0
            seen = []
            for k, (lo, hi, members) in enumerate(buckets):
                assert lo < hi and members == sorted(members, reverse=True)
                if k + 1 < len(buckets):
                    assert buckets[k + 1][1] == lo and hi - lo >= be
                assert lo == offs[members[-1]]
                seen += members
            assert seen == list(range(len(sizes) - 1, -1, -1))


# ---------------------------------------------------------------- device_ops host logic, no GPU needed
def test_basic_index_regions_match_numpy():
    """``_basic_index`` turns a NumPy basic index into (offset, dims, strides) of a C-contiguous array; checked by reading the
    region back through as_strided and comparing with NumPy's own indexing."""
    from numpy.lib.stride_tricks import as_strided
    from compyute_b200.device_ops import _basic_index
    a = np.arange(6 * 5 * 7 * 9, dtype=np.int64).reshape(6, 5, 7, 9)
    keys = [0, -1, (1, 2), (slice(1, 4),), (slice(None), 2), (Ellipsis, 3), (slice(None), slice(None), slice(1, 6, 2), slice(None, None, -1)),
            (slice(4, 1, -1), Ellipsis, slice(2, 3)), (None, 2, Ellipsis), (slice(0, 0),), (2, slice(None), None, 3, slice(8, 2, -3)),
            (Ellipsis,), (5, 4, 6, 8), (slice(10, 20),), (slice(-3, None), slice(None, -2)), (Ellipsis, None), (slice(None, None, -1),) * 4]
    for k in keys:
        off, dims, strides = _basic_index(a.shape, k)
        ref = a[k]
        assert tuple(dims) == ref.shape, (k, dims, ref.shape)
        if ref.size:
            got = as_strided(a.reshape(-1)[off:], shape=dims, strides=[s * 8 for s in strides]) if all(s >= 0 for s in strides) else None
            if got is None:  # negative strides: evaluate element by element
                got = np.empty(dims, np.int64)
                for idx in np.ndindex(*dims):
                    got[idx] = a.reshape(-1)[off + sum(i * s for i, s in zip(idx, strides))]
            assert np.array_equal(got, ref), k
    for bad in [(6,), (0, 0, 0, 0, 0), (Ellipsis, Ellipsis)]:
        with pytest.raises(IndexError):
            _basic_index(a.shape, bad)


def test_broadcast_strides_and_merge():
    from compyute_b200.device_ops import _bstrides, _contig_strides, _merge
    assert _contig_strides((6, 5, 7, 9)) == [315, 63, 9, 1]
    assert _bstrides((5, 1, 1), (6, 5, 7, 9)) == [0, 1, 0, 0]
    assert _bstrides((6, 1, 7, 1), (6, 5, 7, 9)) == [7, 0, 1, 0]
    assert _bstrides((1,), (3, 4)) == [0, 0]
    # same-shape operands collapse to one flat dim; a per-channel operand to (B, C, HW)
    assert _merge((6, 5, 7, 9), [315, 63, 9, 1], [315, 63, 9, 1]) == ([1890], [[1], [1]])
    assert _merge((6, 5, 7, 9), [315, 63, 9, 1], [0, 1, 0, 0]) == ([6, 5, 63], [[315, 63, 1], [0, 1, 0]])
    # every merged form addresses the same elements as the unmerged one
    rng = np.random.RandomState(0)
    for _ in range(50):
        nd = rng.randint(1, 7)
        shape = tuple(int(v) for v in rng.randint(1, 4, nd))
        bshape = tuple(d if rng.rand() < 0.5 else 1 for d in shape)
        sa, sb = _bstrides(shape, shape), _bstrides(bshape, shape)
        md, (ma, mb) = _merge(shape, sa, sb)
        full = [(sum(i * s for i, s in zip(idx, sa)), sum(i * s for i, s in zip(idx, sb))) for idx in np.ndindex(*shape)]
        merged = [(sum(i * s for i, s in zip(idx, ma)), sum(i * s for i, s in zip(idx, mb))) for idx in np.ndindex(*md)] if md else [(0, 0)]
        assert full == merged, (shape, bshape)


def test_tensor_ops_namespace_matches_reference_names():
    """Every function exported by compyute_b200.tensor_ops carries a reference name; on cpu tensors they evaluate with NumPy
    (host staging), which also pins the argument conventions (stop-first arange, dim / keepdims keywords)."""
    import compyute_b200 as cp
    t = cp.tensor(np.arange(12, dtype=np.float32).reshape(3, 4))
    assert np.array_equal(cp.arange(10, 2, 3).to_numpy(), np.arange(2, 10, 3))
    assert np.array_equal(cp.sum(t, 0, keepdims=True).to_numpy(), t.to_numpy().sum(0, keepdims=True))
    assert np.array_equal(cp.concat([t, t], 0).to_numpy(), np.concatenate([t.to_numpy()] * 2, 0))
    assert np.array_equal(cp.pad_to_shape(t, (4, 6)).to_numpy(), np.pad(t.to_numpy(), ((0, 1), (0, 2))))
    assert np.array_equal((t @ t.T).to_numpy(), t.to_numpy() @ t.to_numpy().T)
    assert np.array_equal(cp.maximum(t, 5).to_numpy(), np.maximum(t.to_numpy(), 5))
    assert (None + t).shape == (3, 4) and bool(t == t) is True  # tensors.py:199-201, 305-306
    for name in cp.tensor_ops.__all__:
        assert callable(getattr(cp, name))


# ---------------------------------------------------------------- LR schedulers, gradient clipping (golden = real reference)
def test_lr_schedulers_match_reference_sequences():
    """tests/golden/lr_schedulers.json holds the learning-rate histories the reference's schedulers produce over 30 steps
    (oracle/gen_golden_next.py); ours must reproduce them exactly (same float operations in the same order)."""
    import json
    from compyute_b200.nn import optimizers
    from compyute_b200.nn.utils import lr_schedulers
    cases = json.load(open(os.path.join(GOLDEN, "lr_schedulers.json")))
    assert len(cases) == 5
    for c in cases:
        opt = optimizers.SGD([], lr=c["lr0"])
        sched = getattr(lr_schedulers, c["scheduler"])(opt, **c["kwargs"])
        for i in range(c["steps"]):
            opt.t += 1
            if c["metrics"]:
                sched.step(loss=c["metrics"][i])
            else:
                sched.step()
        assert sched.cache["lr_history"] == c["lr_history"], c["scheduler"]
        assert opt.lr == c["final_lr"]
    with pytest.raises(ValueError):
        lr_schedulers.AdaptiveLrScheduler(optimizers.SGD([], lr=0.1)).step()


def test_clip_grad_norm_host_path_matches_reference():
    import compyute_b200 as cp
    from compyute_b200.nn.parameter import Parameter
    from compyute_b200.nn.utils import clip_grad_norm
    g = np.load(os.path.join(GOLDEN, "clip_grad_norm.npz"))
    for tag in ("loose", "tight"):
        ps = []
        for i in range(5):
            p = Parameter(cp.tensor(np.zeros_like(g[f"g{i}"])))
            p.grad = cp.tensor(g[f"g{i}"].copy())
            ps.append(p)
        norm = clip_grad_norm(iter(ps), float(g[f"{tag}_max_norm"]))
        assert abs(norm - float(g[f"{tag}_norm"])) <= 1e-6 * float(g[f"{tag}_norm"])
        for i, p in enumerate(ps):
            assert np.allclose(p.grad.to_numpy(), g[f"{tag}_g{i}"], rtol=1e-6, atol=1e-8)


# ---------------------------------------------------------------- Sequential peephole planning (no kernels run)
def _host_backed(shape):
    """A DeviceArray handle over HOST memory: enough for the planner predicates, which only look at types, shapes and flags."""
    import torch
    from compyute_b200.tensors import DeviceArray, Tensor
    return Tensor(DeviceArray(torch.zeros(shape), shape, np.float32))


def test_sequential_fusion_planning():
    """Which peepholes the Sequential walk selects (containers.py): BatchNorm->ReLU, BatchNorm2D->ReLU->MaxPooling2D(2),
    residual tails, Linear->ReLU (bf16 mode only), and the staging / statistics hints — and that retain_values, debug mode,
    unsupported shapes or the global switch turn them off."""
    np.random.seed(0)
    seq = nn.Sequential(nn.Conv2D(3, 8, 3, padding="same"), nn.BatchNorm2D(8), nn.ReLU(), nn.MaxPooling2D(2),
                        nn.Conv2D(8, 8, 3, padding="same", bias=False), nn.BatchNorm2D(8), nn.ReLU(),
                        nn.ResidualConnection(nn.Conv2D(8, 8, 3, padding=1), nn.BatchNorm2D(8)), nn.ReLU(),
                        nn.MaxPooling2D(3), nn.Flatten(), nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 10), nn.ReLU(), nn.Linear(10, 4))
    seq.training()
    x4, x2 = _host_backed((2, 8, 8, 8)), _host_backed((2, 32))
    assert seq._pool_fusable(1, x4) and not seq._pool_fusable(5, x4)            # BN, ReLU, MaxPool(2) / no pool behind
    assert not seq._pool_fusable(1, _host_backed((2, 8, 8, 6)))                  # W % 4 != 0 -> separate layers
    assert not seq._pool_fusable(1, _host_backed((2, 8, 7, 8)))                  # H odd
    assert seq._fusable(1, x4) and seq._fusable(5, x4) and not seq._fusable(0, x4)
    assert seq._residual_fusable(7, x4) and not seq._residual_fusable(4, x4)
    with cp.compute_mode("bf16"):
        assert seq._linear_relu_fusable(11, x2)                                  # Out = 64
        assert not seq._linear_relu_fusable(13, _host_backed((2, 64)))           # Out = 10: not a multiple of 32
        assert not seq._linear_relu_fusable(11, _host_backed((2, 3, 32)))        # 3-D input
    with cp.compute_mode("fp32"):
        assert not seq._linear_relu_fusable(11, x2)                              # the epilogue ReLU is a bf16-mode path
    # host tensors, retain_values, debug mode and the global switch disable every peephole
    assert not seq._pool_fusable(1, cp.tensor(np.zeros((2, 8, 8, 8), np.float32)))
    seq.layers[2].retain_values = True
    assert not seq._pool_fusable(1, x4) and not seq._fusable(1, x4)
    seq.layers[2].retain_values = False
    nn.set_fusion_enabled(False)
    try:
        assert not (seq._pool_fusable(1, x4) or seq._fusable(1, x4) or seq._residual_fusable(7, x4))
    finally:
        nn.set_fusion_enabled(True)
    # hints: a convolution followed by BatchNorm2D sums its statistics in the epilogue; a BatchNorm between convolutions writes
    # the channels-last operand of both neighbours; ReLU between Linear layers writes their bf16 rows
    seq._plan_staging_hints()
    assert seq.layers[0]._emit_stats and seq.layers[4]._emit_stats
    assert not seq.layers[1]._emit_cl_fwd and seq.layers[1]._emit_cl_bwd and seq.layers[1]._emit_cl_bwd_sum    # consumer is the pool
    assert seq.layers[5]._emit_cl_fwd and seq.layers[5]._emit_cl_bwd and not seq.layers[5]._emit_cl_bwd_sum    # conv without bias before it
    assert seq.layers[12]._emit_lp_fwd and seq.layers[12]._emit_lp_bwd and not seq.layers[8]._emit_lp_fwd
