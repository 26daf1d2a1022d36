mkdir -p gpurun_out
python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -q -m gpu -x -k "conv or Conv" 2>&1 | tail -4 | cut -c1-300
for v in 0 1; do
  CPT_TC_2CTA_64=$v python bench.py --steps 6 --warmup 3 > gpurun_out/bench_2cta64_$v.json 2> gpurun_out/bench_2cta64_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_2cta64_$v.json').read().strip().splitlines()[-1])
print('2cta64=$v', d['value'], {k:v['ms_fwd_bwd'] for k,v in d['per_layer'].items()}, {k:(round(v['value']),v['ms_per_step']) for k,v in d['models'].items()}, d['parity_check']['ok'], d['tc_watchdog'])
PY
done
