#!/bin/bash
# Weak-scaling run at N GPUs of one box: the default conv sweep and the model workloads, one JSON line each
# (appended to gpurun_out/scale_N.jsonl).  usage: bash tools/scaling_run.sh N [workloads...]
N=${1:-8}; shift
W=${@:-"conv2d_sweep resnet18 vgg mlp"}
O=gpurun_out/scale_$N.jsonl
mkdir -p gpurun_out; : > $O
for w in $W; do
  if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py"; fi
  EXTRA="--steps 8 --warmup 3"
  [ "$w" = "conv2d_sweep" ] && EXTRA="--steps 10 --warmup 3 --no-extra-modes"
  $CMD --gpus $N --workload $w $EXTRA 2> gpurun_out/scale_${N}_$w.err | tail -1 >> $O
  tail -1 $O | cut -c1-160
done
