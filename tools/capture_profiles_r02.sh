#!/bin/bash
# Round-2 profile capture on ONE B200 (run under gpurun): ncu --set full of the C=512 / C=64 convolution launches (-> the
# DRAM traffic record bench.py reads), launch lists of the sweep layers and of one train step of every model config, the
# achieved-GB/s table of the memory-bound kernels.  Only small artefacts are kept (gpurun_out is limited to 64 MiB).
set -u
R=r02
O=gpurun_out
mkdir -p $O
NCU_C=512 ncu --set full --clock-control none -k regex:"tc_kernel|nchw_to_nhwc" -s 10 -c 5 -o $O/${R}_conv512 python tools/ncu_target.py > $O/ncu512.log 2>&1
python tools/ncu_summary.py $O/${R}_conv512.ncu-rep "$R — ncu --set full, Conv2D C=512 B=256 56x56 3x3 bf16 (tools/ncu_target.py; launches: stage x, fprop, stage dy, dgrad, wgrad)" > $O/${R}_conv512_ncu_summary.md
python tools/ncu_traffic.py $O/${R}_conv512.ncu-rep conv512_fprop "tc_kernel.*0, 0, 256, 1, 1" "profiles/${R}_conv512_ncu_summary.md (ncu --set full, fprop / dgrad launches, mean)" > $O/ncu_traffic.log 2>&1
cp profiles/ncu_traffic.json $O/ncu_traffic.json
rm -f $O/${R}_conv512.ncu-rep
NCU_C=64 ncu --set full --clock-control none -k regex:"tc_kernel|nchw_to_nhwc|strip_conv" -s 10 -c 5 -o $O/${R}_conv64 python tools/ncu_target.py > $O/ncu64.log 2>&1
python tools/ncu_summary.py $O/${R}_conv64.ncu-rep "$R — ncu --set full, Conv2D C=64 B=256 56x56 3x3 bf16" > $O/${R}_conv64_ncu_summary.md
rm -f $O/${R}_conv64.ncu-rep
for w in resnet18 vgg mlp mnist; do
  NCU_WORKLOAD=$w ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l_$w.csv python tools/ncu_model.py > /dev/null 2>&1
  python tools/summarize_launches.py $O/l_$w.csv > $O/${R}_launches_$w.md
done
for C in 64 128 256 512; do
  NCU_C=$C NCU_ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_conv$C.csv python tools/ncu_target.py > /dev/null 2>&1
done
python tools/membound_bench.py > $O/${R}_membound_kernels.jsonl 2> $O/membound.err
du -sh $O
