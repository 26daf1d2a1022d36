#!/bin/bash
# A/B of the Conv2D -> ReLU epilogue fusion on config 1 (the reference's MNIST CNN), one B200: CUDA-graph replay and eager,
# B = 128 (the notebook's batch) and B = 4096 (pass-bound instead of launch-bound).
O=gpurun_out; mkdir -p $O; : > $O/r02_conv_relu_ab.jsonl
for B in 128 4096; do for G in "" "--no-graph"; do for F in "" "--no-conv-relu-fusion"; do
  python bench.py --workload mnist --batch $B --steps 20 --warmup 5 $G $F 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'batch': $B, 'graph': '$G' == '', 'conv_relu_fusion': '$F' == '', 'images_per_s': d['value'], 'ms_per_step': d['ms_per_step'], 'gpu_launches': d.get('gpu_launches')}))" | tee -a $O/r02_conv_relu_ab.jsonl
done; done; done
