# A/B of compile-time variants of libmkf_b200 (csrc/Makefile `variant`): tests with $TESTV, bench with each of $VARIANTS
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for tv in $TESTV; do MKF_LIB_VARIANT=$tv timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done
for v in "" $VARIANTS ""; do
  MKF_LIB_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_exp_$v.json 2> gpurun_out/bench_exp_$v.err; tail -c 300 gpurun_out/bench_exp_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_exp_$v.json').read())
r=d['roofline']; e=r['every_slot_computed']
print('[$v] value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'frac',round(r['frac'],3), {k:round(x,4) for k,x in r['stage_ms'].items()}, 'launches', d['gpu_launches'])"
done
