#!/bin/bash
# Final round-2 capture on ONE B200 (run under gpurun): refresh the DRAM-traffic record of the roofline kernel (bench.py reports
# it only while the kernel sources hash to the captured value), the MNIST launch list (Conv2D -> ReLU pairs now fused), then the
# full GPU suite, smoke(), the default bench line and the reference arm.
set -u
R=r02
O=gpurun_out
mkdir -p $O
NCU_C=512 ncu --set full --clock-control none -k regex:"tc_kernel|nchw_to_nhwc" -s 10 -c 5 -o $O/${R}_conv512 python tools/ncu_target.py > $O/ncu512.log 2>&1
python tools/ncu_summary.py $O/${R}_conv512.ncu-rep "$R — ncu --set full, Conv2D C=512 B=256 56x56 3x3 bf16 (tools/ncu_target.py; launches: stage x, fprop, stage dy, dgrad, wgrad)" > $O/${R}_conv512_ncu_summary.md
python tools/ncu_traffic.py $O/${R}_conv512.ncu-rep conv512_fprop "tc_kernel.*0, 0, 256, 1, 1" "profiles/${R}_conv512_ncu_summary.md (ncu --set full, fprop / dgrad launches, mean)" > $O/ncu_traffic.log 2>&1
cp profiles/ncu_traffic.json $O/ncu_traffic.json
rm -f $O/${R}_conv512.ncu-rep
NCU_C=64 NCU_ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_conv64.csv python tools/ncu_target.py > /dev/null 2>&1
NCU_WORKLOAD=mnist ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l_mnist.csv python tools/ncu_model.py > /dev/null 2>&1
python tools/summarize_launches.py $O/l_mnist.csv > $O/${R}_launches_mnist.md
rm -f $O/l_mnist.csv
python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-300 | tee $O/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
python bench.py > $O/${R}_bench_n1.json 2> $O/bench_final.err; tail -c 600 $O/${R}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2>/dev/null; tail -c 300 $O/${R}_bench_reference_arm.json
