#!/bin/bash
# Round profile capture (run on the GPU box): ncu --set full of the conv kernels at C=512 / C=64, a lighter section set for the
# memory-bound kernels of one ResNet-18 step, launch lists, and the achieved-GB/s table.  Reports are summarised on the box
# (tools/ncu_summary.py) and only the small artefacts are kept, so that gpurun_out stays under its 64 MiB limit.
set -u
R=${ROUND:-r01}
O=gpurun_out
mkdir -p $O
for C in 512 64; do
  NCU_C=$C ncu --set full --clock-control none -k regex:"tc_kernel|nchw_to_nhwc" -s 10 -c 5 -o $O/${R}_conv$C python tools/ncu_target.py > $O/ncu$C.log 2>&1
  python tools/ncu_summary.py $O/${R}_conv$C.ncu-rep "$R — ncu --set full, Conv2D C=$C B=256 56x56 3x3 bf16 (tools/ncu_target.py; launches: stage x, fprop, stage dy, dgrad, wgrad)" > $O/${R}_conv${C}_ncu_summary.md
done
rm -f $O/${R}_conv64.ncu-rep
NCU_WORKLOAD=resnet18 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --clock-control none -k regex:"bn_|relu|maxpool|avgpool|adam|softmax|col2im|im2col_pack|add_inplace|wgrad_reduce|w_fprop|w_dgrad" -c 60 \
  -o $O/${R}_resnet_membound python tools/ncu_model.py > $O/ncu_mb.log 2>&1
python tools/ncu_summary.py $O/${R}_resnet_membound.ncu-rep "$R — memory-bound kernels of one ResNet-18 train step (B=256, bf16 mode), ncu SpeedOfLight/Memory sections" > $O/${R}_resnet_membound_ncu_summary.md
rm -f $O/${R}_resnet_membound.ncu-rep
for w in resnet18 vgg mlp mnist; do
  NCU_WORKLOAD=$w ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l_$w.csv python tools/ncu_model.py > /dev/null 2>&1
done
for C in 64 128 256 512; do
  NCU_C=$C NCU_ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l_$C.csv python tools/ncu_target.py > /dev/null 2>&1
done
python tools/membound_bench.py > $O/membound.jsonl 2> $O/membound.err
du -sh $O; ls -la $O | head -40
