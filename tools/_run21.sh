O=gpurun_out; mkdir -p $O
for v in 0 1; do
CPT_TC_2CTA_64=$v NCU_C=64 NCU_ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l64_$v.csv python tools/ncu_target.py > /dev/null 2>&1
python tools/summarize_launches.py $O/l64_$v.csv | head -14
done
CPT_TC_2CTA_64=1 NCU_C=64 ncu --set full --clock-control none -k regex:"tc_kernel" -s 6 -c 3 -o $O/conv64_2cta python tools/ncu_target.py > $O/ncu64_2cta.log 2>&1
python tools/ncu_summary.py $O/conv64_2cta.ncu-rep "C=64 with cta_group::2 64-column tiles" > $O/conv64_2cta_summary.md
rm -f $O/conv64_2cta.ncu-rep
