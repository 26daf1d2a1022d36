"""bench.py — headline benchmark of the Compyute CNN-training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16|tf32|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], the one the "Conv2D fwd+bwd TFLOP/s" half of the metric is quoted on):
``Conv2D(C, C, 3, padding="same")`` for C in {64, 128, 256, 512}, x = (256, C, 56, 56) fp32 NCHW.  One *step* =
forward + backward (dX, dW, db) of all four layers on one batch + the optimizer step (SGD; in data-parallel runs the
gradient arena is all-reduced first).  ``value`` = algorithmic FLOPs of the step (6·B·Co·Ci·Ho·Wo·K² per layer,
SURVEY §8d) ÷ device time, summed over ranks (weak scaling: every rank has its own batch of 256).

Prints ONE JSON line (rank 0).  See the task contract for the keys; extra keys: ``images_per_s``, ``per_layer``,
``modes`` (the other compute modes measured after the timed region), ``tolerance``.

``--impl reference`` times the CPU oracle port (oracle/compyute_ref.py: the reference's as_strided+einsum algorithm,
NumPy) on a bounded sample of the same workload — the reference itself is pure Python and cannot travel to the GPU
box, so kind = "port".
"""

from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SWEEP = (64, 128, 256, 512)
BATCH, HW, KS = 256, 56, 3
TOL = {"fp32": "allclose rtol=atol=1e-5 vs the NumPy reference (db 1e-4)", "fp32_simt": "allclose rtol=atol=1e-5 vs the NumPy reference (db 1e-4)",
       "tf32": "max-abs err <= 2e-3*max|ref|",
       "bf16": "max-abs err <= 1e-2*max|ref|"}


def layer_flops(C: int, B: int) -> float:
    return 6.0 * B * C * C * HW * HW * KS * KS


def peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "MEASURED_PEAKS.json (measured)"}
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_conv_sample(budget_s: float, reps: int = 1):
    """Times the oracle port (reference algorithm, NumPy) on a bounded sample of the sweep.  Returns
    (tflops, sample description, seconds per rep)."""
    from oracle import compyute_ref as R
    # the einsum path runs at ~0.5 GFLOP/s whatever the shape: size the sample to the budget
    cands = [((64, 2), (128, 1)), ((64, 2),), ((64, 1),)]
    est = lambda s: sum(layer_flops(C, b) for C, b in s) / 0.45e9
    sample = next((s for s in cands if est(s) <= budget_s), cands[-1])
    rng = np.random.RandomState(0)
    data = []
    for C, b in sample:
        k = 1.0 / np.sqrt(C * KS * KS)
        data.append((rng.normal(0, 1, (b, C, HW, HW)).astype(np.float32), rng.uniform(-k, k, (C, C, KS, KS)).astype(np.float32),
                     rng.uniform(-k, k, (C,)).astype(np.float32), rng.uniform(-0.1, 0.1, (b, C, HW, HW)).astype(np.float32)))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for x, w, b, dy in data:
            cache = []
            R.conv2d_forward(cache, x, w, b, 1, 1, 1)
            R.conv2d_backward(cache, dy)
        times.append(time.perf_counter() - t0)
    fl = sum(layer_flops(C, b) for C, b in sample)
    desc = "Conv2D 3x3 same 56x56 fwd+bwd, " + " + ".join(f"C={C} at B={b}" for C, b in sample) + " (oracle port: as_strided + numpy.einsum, single-threaded like the reference)"
    return fl / min(times) / 1e12, desc, times


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = max(5.0, 150.0 / (args.steps + args.warmup))
    tf, desc, _ = cpu_conv_sample(budget, 1)  # warm-up + sizing
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.warmup - 1 if args.warmup > 1 else 0):
        cpu_conv_sample(budget, 1)
    for _ in range(args.steps):
        v, desc, _ = cpu_conv_sample(budget, 1)
        vals.append(v)
    value = float(np.median(vals))
    sample_flops = None
    line = {"impl": "reference", "metric": "conv2d_fwd_bwd_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, args.steps + max(0, args.warmup - 1)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, "n/a (CPU)"),
            "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": 1, "host_cores": os.cpu_count(), "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def workload_config(args, mode: str) -> dict:
    return {"workload": "BASELINE configs[1]: single Conv2D layer fwd+bwd sweep, C_in=C_out in {64,128,256,512}, 3x3 same, 56x56, "
                        "batch 256 per GPU, fp32 NCHW in/out, + SGD step (DP: gradient-arena all-reduce)",
            "batch_per_gpu": BATCH, "channels": list(SWEEP), "compute_mode": mode, "tolerance": TOL.get(mode, "n/a"),
            "l2_policy": "inputs larger than L2 (x, dy >= 205 MB per layer vs 126 MB L2)", "parallelism": f"dp{args.gpus}"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args) -> None:
    import torch

    import compyute_b200 as cp
    from compyute_b200 import _lib, distributed, nn

    L = _lib.lib()  # no fallback: raises if the CUDA library is missing
    if args.strip:
        cp.set_strip_conv_enabled(True)
    if args.no_fusion:
        nn.set_fusion_enabled(False)
    if args.no_conv_relu_fusion:
        nn.set_conv_relu_fusion_enabled(False)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # pin the process (CPUs + preferred memory node) to the GPU's NUMA node before any pinned allocation: the e2e numbers are
    # PCIe / host-memory bound (6.2 GB each way per step and rank)
    numa = distributed.bind_to_gpu_numa_node(local) if not args.no_numa_bind else {"bound": False}
    if world > 1:
        distributed.init("nccl")
    dev = cp.cuda
    rng = np.random.RandomState(1234 + rank)

    def make_layers(mode):
        np.random.seed(0)  # identical weights on every rank
        with cp.use_device(dev):
            layers = [nn.Conv2D(C, C, KS, padding="same") for C in SWEEP]
        for l in layers:
            l.training()
        opt = nn.optimizers.SGD([p for l in layers for p in l.get_parameters()], lr=1e-3)
        opt.fused_dp_step = not args.no_fused_step
        return layers, opt

    # device-resident inputs (value) and pinned host copies (e2e)
    xs = [torch.randn(BATCH, C, HW, HW, device="cuda", generator=torch.Generator("cuda").manual_seed(10 + rank)) for C in SWEEP]
    dys = [torch.empty(BATCH, C, HW, HW, device="cuda").uniform_(-0.1, 0.1) for C in SWEEP]
    from compyute_b200.tensors import DeviceArray, Tensor
    wrap = lambda t: Tensor(DeviceArray(t, tuple(t.shape), np.float32))
    x_t, dy_t = [wrap(t) for t in xs], [wrap(t) for t in dys]

    probe = {"ev": [], "C": 512}
    orig_fprop = L.cpt_conv2d_fprop_cl

    def fprop_probe(d, *a):  # CUDA events on the launching stream around the dominant kernel's C-ABI call
        dd = d._obj if hasattr(d, "_obj") else d
        if probe.get("on") and dd.Ci == probe["C"]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = orig_fprop(d, *a); e1.record()
            probe["ev"].append((e0, e1))
            return r
        return orig_fprop(d, *a)

    def step(layers, opt):
        opt.reset_grads()
        for l, x, dy in zip(layers, x_t, dy_t):
            l(x)
            l.backward(dy)
        opt.step()

    def timed(layers, opt, steps, warmup, with_probe=False):
        for _ in range(warmup):
            step(layers, opt)
        torch.cuda.synchronize()
        if world > 1:
            distributed.barrier()
        probe["on"] = with_probe
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.cpt_launch_count()
        gc.collect(); gc.disable()  # a cyclic-GC pause on the host starves the launch queue (seen as one stalled step)
        try:
            e0.record()
            for _ in range(steps):
                step(layers, opt)
            e1.record()
            torch.cuda.synchronize()
        finally:
            gc.enable()
        probe["on"] = False
        if world > 1:
            distributed.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, L.cpt_launch_count() - n0

    step_flops = sum(layer_flops(C, BATCH) for C in SWEEP)
    pk = peaks()
    sampler = ClockSampler(local)
    results = {}
    with cp.compute_mode(args.mode):
        L.cpt_conv2d_fprop_cl = fprop_probe if args.mode != "fp32_simt" else orig_fprop
        layers, opt = make_layers(args.mode)
        if rank == 0:
            sampler.start()
        ms, launches = timed(layers, opt, args.steps, args.warmup, with_probe=True)
        clocks = sampler.stop() if rank == 0 else {}
        grad_sync = "n/a" if world == 1 else (("fused sharded step: " + opt.fused_dp_note()) if opt._fused is not None else
                                              "one NCCL all-reduce of the gradient arena at step() + replicated update (" + opt.fused_dp_note() + ")")
        L.cpt_conv2d_fprop_cl = orig_fprop
        status = L.cpt_tc_check_status()
        # per-layer device times (separate short pass, for the report)
        per_layer = {}
        for l, x, dy, C in zip(layers, x_t, dy_t, SWEEP):
            def one():
                l(x); l.backward(dy)
            for _ in range(2):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                one()
            e1.record(); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 3
            per_layer[f"C{C}"] = {"ms_fwd_bwd": round(t, 4), "tflops": round(layer_flops(C, BATCH) / t / 1e9, 1)}

        # ---- end to end through the public API with HOST buffers (pinned): H2D of x, w, b, dy; D2H of y, dx, dw, db.
        # Copies run on their own streams so that the H2D of layer i+1 and the D2H of layer i-1 overlap layer i's kernels
        # (PCIe is full duplex); everything is inside the timed region and the step ends with a full synchronize.
        e2e = None
        if rank == 0 or world > 1:
            from compyute_b200.nn.functional import Conv2DFn, FunctionCache
            e2e_steps = max(1, min(args.steps, 3))
            Cmax = max(SWEEP)
            pin = lambda n: torch.empty(n, dtype=torch.float32).pin_memory()
            hx, hdy = pin(BATCH * Cmax * HW * HW), pin(BATCH * Cmax * HW * HW)
            hx.normal_(); hdy.uniform_(-0.1, 0.1)
            hw, hb = pin(Cmax * Cmax * KS * KS), pin(Cmax)
            hw.uniform_(-0.01, 0.01); hb.uniform_(-0.01, 0.01)
            bufs = {}
            # one pinned landing buffer per result kind, sized for the largest layer and reused by all four layers (the D2H
            # stream is serial, so reuse costs nothing and keeps the pinned footprint at 6.6 GB per rank)
            hy_all, hdx_all, hdw_all, hdb_all = pin(BATCH * Cmax * HW * HW), pin(BATCH * Cmax * HW * HW), pin(Cmax * Cmax * KS * KS), pin(Cmax)
            for C in SWEEP:
                n, nw = BATCH * C * HW * HW, C * C * KS * KS
                bufs[C] = dict(n=n, nw=nw, x=torch.empty(BATCH, C, HW, HW, device="cuda"), dy=torch.empty(BATCH, C, HW, HW, device="cuda"),
                               w=torch.empty(C, C, KS, KS, device="cuda"), b=torch.empty(C, device="cuda"),
                               hy=hy_all[:n], hdx=hdx_all[:n], hdw=hdw_all[:nw], hdb=hdb_all[:C], free=torch.cuda.Event(),
                               ready=torch.cuda.Event(), ready_dy=torch.cuda.Event(), done_fwd=torch.cuda.Event(),
                               done=torch.cuda.Event(), keep=None)
            s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
            h2d = d2h = 0

            def e2e_step(count=False):
                nonlocal h2d, d2h
                for C in SWEEP:
                    q = bufs[C]
                    with torch.cuda.stream(s_in):
                        s_in.wait_event(q["free"])  # previous step's kernels no longer read these device buffers
                        q["x"].copy_(hx[:q["n"]].view_as(q["x"]), non_blocking=True)
                        q["w"].copy_(hw[:q["nw"]].view_as(q["w"]), non_blocking=True)
                        q["b"].copy_(hb[:C], non_blocking=True)
                        q["ready"].record(s_in)       # forward can start while dy is still in flight
                        q["dy"].copy_(hdy[:q["n"]].view_as(q["dy"]), non_blocking=True)
                        q["ready_dy"].record(s_in)
                    s_cmp.wait_event(q["ready"])
                    c = FunctionCache()
                    y = Conv2DFn.forward(c, wrap(q["x"]), wrap(q["w"]), wrap(q["b"]), 1, 1, 1)
                    q["done_fwd"].record(s_cmp)
                    with torch.cuda.stream(s_out):    # y goes home while backward runs
                        s_out.wait_event(q["done_fwd"])
                        y.data._buf.record_stream(s_out)
                        q["hy"].copy_(y.data._buf.view(-1), non_blocking=True)
                    s_cmp.wait_event(q["ready_dy"])
                    gx, gw, gb = Conv2DFn.backward(c, wrap(q["dy"]))
                    q["free"].record(s_cmp); q["done"].record(s_cmp)
                    outs = (gx.data._buf, gw.data._buf, gb.data._buf)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(q["done"])
                        for t_dev, t_host in zip(outs, (q["hdx"], q["hdw"], q["hdb"])):
                            t_dev.record_stream(s_out)
                            t_host.copy_(t_dev.view(-1), non_blocking=True)
                    if count:
                        h2d += 4 * (2 * q["n"] + q["nw"] + C); d2h += 4 * (2 * q["n"] + q["nw"] + C)
                torch.cuda.synchronize()

            e2e_step()
            if world > 1:
                distributed.barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                e2e_step(count=(i == 0))
            dt = (time.perf_counter() - t0) / e2e_steps
            if world > 1:
                t = torch.tensor([dt], device="cuda")
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                dt = float(t.item())
            e2e = {"value": round(world * step_flops / dt / 1e12, 3), "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "steps": e2e_steps, "ms_per_step": round(dt * 1e3, 2),
                   "host_link_gbs_per_rank_each_way": round(h2d / dt / 1e9, 1), "host_link_gbs_all_ranks_both_ways": round(world * (h2d + d2h) / dt / 1e9, 1),
                   "numa": numa,
                   "note": "Conv2DFn.forward/backward per layer with pinned host buffers: H2D x,w,b,dy and D2H y,dx,dw,db inside the timed "
                           "region, copies on side streams overlapping the kernels (forward starts when x has landed, y returns while backward "
                           "runs); PCIe-bound (6.2 GB each way per step)"}
            del hx, hdy, bufs

    # ---- parity of what was just timed (outside the timed region): the C=512 layer's y / dx / dw of one more pass in the
    # benchmarked mode against fp64 dot products of sampled entries (operands fetched from the device tensors of this run)
    parity = None
    if rank == 0:
        with cp.compute_mode(args.mode):
            parity = parity_check(layers[-1], xs[-1], dys[-1], x_t[-1], dy_t[-1], args.mode)

    # ---- other compute modes (outside the headline timed region; same step, fewer iterations)
    modes = {}
    if rank == 0 and world == 1 and not args.no_extra_modes:
        for m, (k, w) in {"bf16": (3, 2), "tf32": (3, 2), "fp32": (2, 1), "fp32_simt": (1, 1)}.items():
            if m == args.mode:
                continue
            with cp.compute_mode(m):
                lay, op = make_layers(m)
                mms, _ = timed(lay, op, k, w)
            modes[m] = {"tflops": round(step_flops / mms / 1e9, 1), "ms_per_step": round(mms, 3), "tolerance": TOL[m]}

    # ---- the "CNN train images/s" half of the metric: configs 2-4 through the module API at this world size, a few steps
    # each (outside the headline timed region; full records: --workload vgg|resnet18|mlp)
    models = {}
    if not args.no_models:
        del layers, opt
        xs.clear(); dys.clear(); x_t.clear(); dy_t.clear()
        gc.collect(); torch.cuda.empty_cache()
        plan = [("resnet18", False), ("vgg", False), ("mlp", False)] + ([("vgg", True)] if world > 1 else [])
        for wl, strong in plan:
            sub = argparse.Namespace(**vars(args))
            sub.workload, sub.batch, sub.strong, sub.steps, sub.warmup = wl, 0, strong, args.model_steps, 3
            sub.graph = False
            sub.overlap = False
            try:
                rec = model_record(sub, cpu=False)
            except Exception as e:  # pragma: no cover - reported, never silent
                rec = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
            gc.collect(); torch.cuda.empty_cache()
            if rec is not None:
                keep = ("value", "unit", "ms_per_step", "scaling", "tflops", "gpu_launches", "steps", "tc_watchdog", "error")
                slim = {k: rec[k] for k in keep if k in rec}
                if "config" in rec:
                    slim.update(batch_per_gpu=rec["config"]["batch_per_gpu"], global_batch=rec["config"]["global_batch"],
                                cuda_graph=rec["config"]["cuda_graph"], grad_sync=rec["config"]["grad_sync"])
                    slim["e2e"] = rec["e2e"]["value"]
                    slim["frac_of_bf16_peak"] = rec["roofline"]["frac"]
                models[wl + ("_strong" if strong else "")] = slim
    if rank != 0:
        return
    # dominant kernel: tc_kernel<bf16, OP_CONV, BN=256> for the C=512 fprop (same kernel runs dgrad)
    roof = None
    if probe["ev"]:
        kms = float(np.mean([a.elapsed_time(b) for a, b in probe["ev"]]))
        fl = layer_flops(512, BATCH) / 3.0
        ach = fl / kms / 1e9
        roof = {"bound": "tensor", "kernel": "tc_kernel<BF16,K-major,K-major,BN=256,OP_CONV> (C=512 fprop; includes the <1% weight re-layout launch)",
                "achieved": round(ach, 1), "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": round(ach / pk["bf16_tflops"], 4),
                "frac_of_sustained": round(ach / pk["bf16_tflops_sustained"], 4) if pk.get("bf16_tflops_sustained") else None,
                "peak_source": pk["source"] + " (burst cuBLAS bf16)", "ms_per_launch": round(kms, 4), "launches_timed": len(probe["ev"]),
                "flops_per_launch": fl,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture of THIS
                # build of the kernel sources (profiles/ncu_traffic.json, keyed by a hash of tc_kernel.cuh + tc_host.cu);
                # null when the sources changed since the capture.  Algorithmic bytes = x_cl bf16 0.82 GB + y fp32 1.64 GB + filters
                "algorithmic_bytes": 2.47e9, **ncu_traffic("conv512_fprop")}
    cpu_tf, cpu_desc, cpu_times = cpu_conv_sample(12.0, 1)
    value = world * step_flops / ms / 1e9
    line = {"metric": "conv2d_fwd_bwd_tflops", "value": round(value, 2), "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32", "fp32_simt": "f32"}[args.mode], "data": "synthetic",
            "config": dict(workload_config(args, args.mode), grad_sync=grad_sync), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": {"value": round(cpu_tf, 7), "unit": "TFLOP/s", "cores": 1, "host_cores": os.cpu_count(),
                                                "kind": "port", "sample": cpu_desc, "seconds": round(cpu_times[0], 2)},
            "images_per_s": round(world * BATCH * len(SWEEP) / (ms / 1e3), 1), "frac_of_bf16_peak": round(value / world / pk["bf16_tflops"], 4),
            "per_layer": per_layer, "modes": modes, "models": models, "parity_check": parity, "tc_watchdog": int(status)}
    emit(line)


def kernel_source_hash() -> str:
    import hashlib
    h = hashlib.sha1()
    for f in ("tc_kernel.cuh", "tc_host.cu", "tc_ptx.cuh"):
        with open(os.path.join(ROOT, "compyute_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def ncu_traffic(key: str) -> dict:
    """DRAM bytes per launch of a kernel from the committed ncu capture of the current kernel sources (never a constant)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            db = json.load(f)
        rec = db.get(key, {})
        if rec.get("source_hash") == kernel_source_hash():
            return {"traffic": rec["dram_bytes"], "traffic_source": rec.get("capture", "profiles/ncu_traffic.json")}
        return {"traffic": None, "traffic_source": f"no ncu capture for kernel sources {kernel_source_hash()} (last: {rec.get('source_hash')}, {rec.get('dram_bytes')} B)"}
    except Exception as e:
        return {"traffic": None, "traffic_source": f"unavailable: {e}"}


def parity_check(layer, x_dev, dy_dev, x_t, dy_t, mode: str, samples: int = 96) -> dict:
    """One more forward + backward of the C=512 layer in the benchmarked mode; sampled entries of y, dx and dw against
    fp64 dot products.  Errors are relative to the largest magnitude of the tensor (the mode's tolerance convention)."""
    import torch
    tol = {"bf16": 1e-2, "tf32": 2e-3, "fp32": 1e-5, "fp32_simt": 1e-5}[mode]
    layer.w.grad = None  # first gradient of a step is stored by reference (module.py:392-400), later ones accumulate
    if layer.b is not None:
        layer.b.grad = None
    y = layer(x_t)
    dx = layer.backward(dy_t)
    yb, dxb = y.data._buf, dx.data._buf
    B, C, H, _ = x_dev.shape
    w = layer.w.data._buf.reshape(C, C, 3, 3).double()  # a flat slice of the symmetric parameter arena in fused data-parallel runs
    b = layer.b.data._buf.reshape(C).double()
    dw = layer.w.grad.data._buf.reshape(C, C, 3, 3)  # a flat slice of the optimizer's gradient arena
    g = torch.Generator().manual_seed(0)
    r = lambda n: int(torch.randint(0, n, (1,), generator=g))
    xp = lambda t, bi, p, q: torch.nn.functional.pad(t[bi].double(), (1, 1, 1, 1))[:, p:p + 3, q:q + 3]
    ey = edx = edw = 0.0
    for _ in range(samples):
        bi, o, p, q = r(B), r(C), r(H), r(H)
        ref = (xp(x_dev, bi, p, q) * w[o]).sum() + b[o]
        ey = max(ey, abs(float(yb[bi, o, p, q]) - float(ref)))
        ref = (xp(dy_dev, bi, p, q) * w[:, o].flip(-1, -2)).sum()
        edx = max(edx, abs(float(dxb[bi, o, p, q]) - float(ref)))
    for _ in range(16):
        o, i, j, k = r(C), r(C), r(3), r(3)
        xs = torch.nn.functional.pad(x_dev[:, i].double(), (1, 1, 1, 1))[:, j:j + H, k:k + H]
        ref = (dy_dev[:, o].double() * xs).sum()
        edw = max(edw, abs(float(dw[o, i, j, k]) - float(ref)))
    sy, sdx, sdw = float(yb.abs().max()), float(dxb.abs().max()), float(dw.abs().max())
    rel = {"y": ey / sy, "dx": edx / sdx, "dw": edw / sdw}
    return {"layer": "Conv2D(512, 512, 3, same), x = (256, 512, 56, 56)", "mode": mode, "samples": {"y": samples, "dx": samples, "dw": 16},
            "max_err_over_max_abs": {k: float(f"{v:.3e}") for k, v in rel.items()}, "tolerance": tol,
            "reference": "fp64 dot products (torch, on the device tensors of the run)", "ok": bool(all(v <= tol for v in rel.values()))}


# ------------------------------------------------------------------------------------------------ model workloads
MODEL_WORKLOADS = {
    # name: (spec factory, input shape per sample, classes, batch per GPU, BASELINE config text)
    "mnist": ("mnist_cnn", (1, 28, 28), 10, 128, "BASELINE configs[0]: example_cnn_mnist CNN, synthetic 1x28x28, batch 128/GPU, Adam, CE"),
    "vgg": ("vgg", (3, 32, 32), 10, 1024, "BASELINE configs[2]: VGG-style Conv2D+BatchNorm2D+MaxPool2D CNN, synthetic 3x32x32, batch 1024/GPU, Adam, CE"),
    "resnet18": ("resnet18", (3, 224, 224), 1000, 256, "BASELINE configs[3]: ResNet-18-shaped Sequential CNN, synthetic 3x224x224, batch 256/GPU, Adam, CE"),
    "mlp": ("mlp", (4096,), 4096, 8192, "BASELINE configs[4]: MLP 8 x Linear(4096,4096), batch 8192/GPU, Adam, CE"),
}


def run_model(args) -> None:
    line = model_record(args)
    if line is not None:
        emit(line)


def model_record(args, cpu: bool = True):
    """Train-step throughput (images/s) of one of the model configs through the public module API.  Returns the JSON
    record on rank 0 (None elsewhere)."""
    import torch

    import bench_workloads as W
    import compyute_b200 as cp
    from compyute_b200 import _lib, distributed, nn
    from compyute_b200.tensors import DeviceArray, Tensor

    L = _lib.lib()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not getattr(args, "no_numa_bind", False) and not getattr(model_record, "_bound", False):
        distributed.bind_to_gpu_numa_node(local)
        model_record._bound = True
    if world > 1 and not distributed.is_initialized():
        distributed.init("nccl")
    if getattr(args, "no_fusion", False):
        nn.set_fusion_enabled(False)
    if getattr(args, "no_conv_relu_fusion", False):
        nn.set_conv_relu_fusion_enabled(False)
    factory, xshape, classes, B, desc = MODEL_WORKLOADS[args.workload]
    B = args.batch or B
    scaling = "weak"
    if getattr(args, "strong", False):  # strong scaling: the GLOBAL batch is fixed, every rank takes its shard
        assert B % world == 0, "strong scaling needs the global batch to divide by the world size"
        B //= world
        scaling = "strong"
    spec = getattr(W, factory)()
    hw = xshape[-1] if len(xshape) == 3 else 1
    flops_img = W.train_flops_per_image(spec, hw)
    np.random.seed(0)
    with cp.use_device(cp.cuda), cp.compute_mode(args.mode):
        model = W.build(spec)
    model.training()
    opt = nn.optimizers.Adam(model.get_parameters(), lr=1e-3)
    opt.overlap_grad_sync = world > 1 and args.overlap  # bucketed all-reduces launched during backward
    opt.fused_dp_step = world > 1 and not args.no_fused_step  # exchange + update in one kernel over NVLS / peer memory
    distributed.set_sync_batchnorm(world > 1 and args.sync_bn)  # BatchNorm statistics over the global batch (optional)
    opt.reserve_sms = int(os.environ.get("CPT_DP_RESERVE_SMS", opt.reserve_sms))
    opt.bucket_bytes = int(os.environ.get("CPT_DP_BUCKET_BYTES", opt.bucket_bytes))
    loss_fn = nn.CrossEntropyLoss()
    g = torch.Generator("cuda").manual_seed(100 + rank)
    wrapf = lambda t: Tensor(DeviceArray(t, tuple(t.shape), np.float32))
    wrapi = lambda t: Tensor(DeviceArray(t, tuple(t.shape), np.int32))
    hx = torch.randn(B, *xshape).pin_memory()
    ht = torch.randint(0, classes, (B,), dtype=torch.int32).pin_memory()
    dx, dt = hx.cuda(), ht.cuda()
    hloss = torch.zeros(1).pin_memory()

    def step(x, t):
        loss = loss_fn(model(x), t)
        opt.reset_grads()
        model.backward(loss_fn.backward())
        opt.step()
        return loss

    def step_resident():
        return step(wrapf(dx), wrapi(dt))

    # e2e: every step copies one batch host->device from pinned memory and reads the loss back.  The copy of step i+1's batch is
    # issued on a side stream before step i's kernels (double-buffered device inputs), so PCIe overlaps compute the way the
    # Dataloader's pinned staging does; the host waits for the loss of step i only (event on the compute stream).
    s_in = torch.cuda.Stream()
    s_cmp = torch.cuda.current_stream()
    ebuf = [dict(x=torch.empty_like(dx), t=torch.empty_like(dt), ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    estate = {"i": 0, "primed": False}
    loss_ev = torch.cuda.Event()

    def _issue(k):
        q = ebuf[k]
        with torch.cuda.stream(s_in):
            s_in.wait_event(q["free"])  # the step that read this buffer has finished
            q["x"].copy_(hx, non_blocking=True)
            q["t"].copy_(ht, non_blocking=True)
            q["ready"].record(s_in)

    def step_e2e():
        k = estate["i"] & 1
        estate["i"] += 1
        if not estate["primed"]:
            _issue(k)
            estate["primed"] = True
        _issue(k ^ 1)  # next step's batch travels while this step computes
        q = ebuf[k]
        s_cmp.wait_event(q["ready"])
        loss = step(wrapf(q["x"]), wrapi(q["t"]))
        q["free"].record(s_cmp)
        hloss.copy_(loss.data._buf.view(1), non_blocking=True)
        loss_ev.record(s_cmp)
        loss_ev.synchronize()

    def timed(fn, steps, warmup, device_timed=True):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            distributed.barrier()
        n0 = L.cpt_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect(); gc.disable()  # a cyclic-GC pause on the host starves the launch queue of the small-image workloads
        try:
            t0 = time.perf_counter(); e0.record()
            for _ in range(steps):
                fn()
            e1.record(); torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
        finally:
            gc.enable()
        if world > 1:
            distributed.barrier()
        ms = e0.elapsed_time(e1) if device_timed else wall
        if world > 1:
            tt = torch.tensor([ms], device="cuda"); torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX); ms = float(tt.item())
        return ms / steps, L.cpt_launch_count() - n0

    graphed = None
    if args.workload in ("mnist", "vgg") and world == 1 and not args.no_graph:
        # configs 1 and 2 are launch-bound on one GPU (mnist: 645 launches of a few microseconds per step; vgg: 1000 launches in
        # 4.5 ms, 226-231 k images/s eager vs 246 k replayed): the step is replayed as one CUDA graph
        args.graph = True
    if world > 1 and args.graph_dp and args.workload in ("mnist", "vgg") and not args.no_graph and not args.overlap:
        # data-parallel CUDA graph: needs the fused sharded step (its barriers and kernel are plain launches; an NCCL call would
        # be captured too, but only the fused form is validated).  One eager step first: it moves the arenas into symmetric memory.
        with cp.compute_mode(args.mode):
            step_resident()
        args.graph = opt._fused is not None
    if args.graph and (world == 1 or opt._fused is not None):
        # CUDA-graph replay of the whole step (fwd + loss + bwd + fused Adam): static input tensors, one launch per step
        xs_t, ts_t = wrapf(dx), wrapi(dt)
        with cp.compute_mode(args.mode):
            graphed = cp.graph.CapturedStep(lambda: step(xs_t, ts_t), optimizers=[opt], warmup=3)
        step_resident = graphed  # noqa: F811

        def step_e2e():  # noqa: F811
            dx.copy_(hx, non_blocking=True); dt.copy_(ht, non_blocking=True)
            loss = graphed()
            hloss.copy_(loss.data._buf.view(1), non_blocking=True)
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    with cp.compute_mode(args.mode):
        if rank == 0:
            sampler.start()
        ms, launches = timed(step_resident, args.steps, args.warmup)
        clocks = sampler.stop() if rank == 0 else {}
        e2e_ms, _ = timed(step_e2e, max(2, args.steps // 2), 2, device_timed=False)
    status = L.cpt_tc_check_status()
    if rank != 0:
        return None
    pk = peaks()
    ips = world * B / (ms / 1e3)
    tfl = ips * flops_img / 1e12
    # CPU oracle port on a bounded sample: a few images through the same spec (fwd + bwd + Adam)
    from oracle.model_ref import RefModel
    nb = {"mnist": 16, "vgg": 2, "resnet18": 1, "mlp": 64}[args.workload]
    if args.workload == "resnet18":  # a full 224x224 ResNet-18 image costs ~5 h on the einsum path: time the stem stage only
        cpu_spec = spec[:4]
        cpu_note = "stem only (Conv7x7/s2+BN+ReLU+MaxPool), 1 image; the full model is ~10.9 GFLOP/image at ~1 GFLOP/s"
    else:
        cpu_spec, cpu_note = spec, f"{nb} images, fwd + bwd"
    cpu_val = None
    try:
        if not cpu:
            raise RuntimeError("skipped (sub-record of the default line; run --workload for the CPU sample)")
        rs = np.random.RandomState(0)
        def rand_params(sp):
            ps, bs = [], []
            for s_ in sp:
                if s_[0] == "conv":
                    ps.append(rs.normal(0, 0.05, (s_[2], s_[1], s_[3], s_[3])).astype(np.float32))
                    if s_[6]: ps.append(np.zeros(s_[2], np.float32))
                elif s_[0] == "linear":
                    ps.append(rs.normal(0, 0.02, (s_[2], s_[1])).astype(np.float32))
                    if s_[3]: ps.append(np.zeros(s_[2], np.float32))
                elif s_[0] in ("bn2d", "bn1d"):
                    ps += [np.ones(s_[1], np.float32), np.zeros(s_[1], np.float32)]; bs += [np.zeros(s_[1], np.float32), np.ones(s_[1], np.float32)]
                elif s_[0] == "residual":
                    for sub in (s_[1], s_[2] or []):
                        p2, b2 = rand_params(sub); ps += p2; bs += b2
            return ps, bs
        ps, bs = rand_params(cpu_spec)
        ref = RefModel(cpu_spec, ps, bs)
        xs = rs.normal(0, 1, (nb, *xshape)).astype(np.float32)
        t0 = time.perf_counter()
        y = ref.forward(xs, True)
        ref.backward(np.ones_like(y) / y.size)
        cpu_s = time.perf_counter() - t0
        cpu_val = nb / cpu_s
        if args.workload == "resnet18":
            cpu_val = cpu_val * (W.train_flops_per_image(cpu_spec, hw) / flops_img)  # images/s scaled by the FLOP share of the timed stage
            cpu_note += " — value = measured stem images/s x (stem FLOPs / model FLOPs), i.e. an optimistic bound for the CPU"
    except Exception as e:  # pragma: no cover
        cpu_note = f"cpu sample failed: {e}"
    line = {"metric": "cnn_train_images_per_s" if args.workload != "mlp" else "mlp_train_samples_per_s", "value": round(ips, 1), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32", "fp32_simt": "f32"}[args.mode], "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "global_batch": B * world, "compute_mode": args.mode, "cuda_graph": bool(graphed), "tolerance": TOL[args.mode], "parallelism": f"dp{world}",
                       "grad_sync": (("fused sharded step: " + opt.fused_dp_note()) if opt._fused is not None else
                                     ("bucketed all-reduce overlapped with backward" if opt.overlap_grad_sync else "one NCCL all-reduce at step() + replicated update")) if world > 1 else "n/a",
                       "batchnorm": "synchronised (global-batch statistics)" if distributed.sync_batchnorm_active() else "per-shard statistics",
                       "l2_policy": "activations of one step exceed L2" if B * int(np.prod(xshape)) * 4 > 126e6 else "L2 flushed implicitly: per-step activation traffic exceeds L2"},
            "clocks": clocks, "e2e": {"value": round(world * B / (e2e_ms / 1e3), 1), "unit": "images/s", "h2d_bytes_per_step": int(hx.numel() * 4 + ht.numel() * 4),
                                      "d2h_bytes_per_step": 4, "note": "module API; one batch H2D from pinned memory (prefetched on a copy stream during the previous step) and the loss D2H every step; wall clock incl. host dispatch"},
            "gpu_launches": int(launches), "roofline": {"bound": "tensor", "achieved": round(tfl / world, 2), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                                                        "frac": round(tfl / world / pk["bf16_tflops"], 4), "traffic": None,
                                                        "note": f"whole step: {flops_img / 1e9:.3f} GFLOP/image algorithmic (3 x contraction FLOPs)"},
            "cpu_baseline": {"value": None if cpu_val is None else round(cpu_val, 4), "unit": "images/s", "cores": 1, "host_cores": os.cpu_count(), "kind": "port", "sample": cpu_note},
            "tflops": round(tfl, 2), "tc_watchdog": int(status)}
    return line


_REAL_STDOUT = None


def _protect_stdout() -> None:
    """Everything that writes to fd 1 while the benchmark runs (NCCL prints its version banner there) is sent to stderr;
    the ONE JSON line is written to the original stdout by ``emit``."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.__stdout__
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("COMPYUTE_B200_MODE", "bf16"), choices=["bf16", "tf32", "fp32", "fp32_simt"])
    ap.add_argument("--no-extra-modes", action="store_true")
    ap.add_argument("--no-models", action="store_true", help="skip the model sub-records (configs 2-4) of the default line")
    ap.add_argument("--model-steps", type=int, default=5, help="timed steps of each model sub-record")
    ap.add_argument("--strong", action="store_true", help="model workloads: strong scaling (the configured batch is the GLOBAL batch)")
    ap.add_argument("--workload", default="conv2d_sweep", choices=["conv2d_sweep", "mnist", "vgg", "resnet18", "mlp"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of a model workload")
    ap.add_argument("--graph", action="store_true", help="model workloads: replay the train step as one CUDA graph (default for mnist and vgg on one GPU)")
    ap.add_argument("--graph-dp", action="store_true", help="mnist / vgg at N > 1: replay the data-parallel step (fused sharded step included) as one CUDA graph.  Opt-in: measured "
                         "faster at 2 GPUs (4.28 vs 4.62 ms VGG) and for the launch-bound strong-scaling case at 8 (1.46 vs 3.69 ms), slower for "
                         "weak-scaling VGG at 8 GPUs (6.7 vs 4.6 ms; not diagnosed)")
    ap.add_argument("--no-graph", action="store_true", help="mnist / vgg: eager launches instead of the default CUDA-graph replay")
    ap.add_argument("--overlap", action="store_true",
                    help="data-parallel model runs: bucketed all-reduces launched during backward (Optimizer.overlap_grad_sync) "
                         "instead of one all-reduce of the whole gradient arena at step(); measured gain at 2 GPUs is ~1 %% because "
                         "the persistent GEMM grids leave NCCL little room, so it is opt-in")
    ap.add_argument("--strip", action="store_true", help="enable the opt-in strip (shared-halo) convolution kernels for the C <= 128 layers")
    ap.add_argument("--no-fusion", action="store_true", help="model workloads: evaluate layer by layer (no Sequential peephole fusions), for A/B runs")
    ap.add_argument("--no-conv-relu-fusion", action="store_true", help="model workloads: Conv2D -> ReLU pairs as two layers (A/B of that one fusion)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA node")
    ap.add_argument("--no-fused-step", action="store_true",
                    help="data-parallel runs: classic NCCL all-reduce + replicated update instead of the fused sharded step over "
                         "NVLS / peer memory (csrc/dp_step.cu)")
    ap.add_argument("--sync-bn", action="store_true",
                    help="data-parallel model runs: synchronised BatchNorm (distributed.set_sync_batchnorm): statistics and backward "
                         "sums over the global batch, one small all-gather / all-reduce per BatchNorm layer and pass")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "conv2d_sweep":
        run_model(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
