"""Data-parallel correctness on real GPUs (``-m gpu``; needs >= 2 visible devices, skips otherwise): ``tools/dp_check.py``
under ``torch.distributed.run`` — every rank trains on its shard, rank 0 also trains the full batch alone, and parameters,
running statistics and the loss trace must agree (1e-5 in fp32 mode; the replicas must be bit-identical among themselves).

Variants: the fused sharded step over NVLS / peer memory (csrc/dp_step.cu, the default), the classic exchange through the
library's own NCCL communicator (``cpt_nccl_*``) + replicated update, the overlapped bucketed exchange, and synchronised
BatchNorm.  The host-side logic of all of them (sharding, bucket planning, the SyncBN merge, shard layout of the fused
step) is covered on CPU with gloo at world size 2 in tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "fused_step": {},                                   # peer loads / stores at 2 ranks, multimem from 3 ranks on
    "fused_step_multimem": {"CPT_DP_MULTIMEM": "1"},    # multimem.ld_reduce / multimem.st through the NVSwitch at any world size
    "fused_step_peer": {"CPT_DP_MULTIMEM": "0"},
    "fused_step_overlapped_buckets": {"DP_OVERLAP": "1"},                       # bucketed, on a side stream during backward
    "fused_step_overlapped_multimem": {"DP_OVERLAP": "1", "CPT_DP_MULTIMEM": "1"},
    "nccl_allreduce": {"DP_FUSED": "0"},
    "nccl_overlapped_buckets": {"DP_FUSED": "0", "DP_OVERLAP": "1"},
    "torch_distributed_allreduce": {"DP_FUSED": "0", "CPT_OWN_NCCL": "0"},
    "clip_grad_norm_fused": {"DP_CLIP": "0.05"},                                 # sync_grads() first, the fused kernel then skips its sum
    "clip_grad_norm_nccl": {"DP_CLIP": "0.05", "DP_FUSED": "0"},
    "sync_batchnorm": {"DP_SYNCBN": "1"},
    "fused_step_bf16": {"DP_MODE": "bf16", "DP_SYNCBN": "1"},
}


@pytest.mark.parametrize("name", list(VARIANTS), ids=list(VARIANTS))
def test_data_parallel_equals_single_process(name):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, found {n}")
    world = 2 if n < 4 else 4
    env = dict(os.environ, **VARIANTS[name])
    port = 29600 + (os.getpid() + list(VARIANTS).index(name)) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = (r.stdout or "") + (r.stderr or "")
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, f"dp_check_{name}.log"), "w") as f:
            f.write(out)
    assert r.returncode == 0 and "DP_CHECK OK" in out, out[-3000:]
