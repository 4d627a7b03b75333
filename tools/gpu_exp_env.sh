# A/B of run-time knobs: bench.py once per value of $VAR in $VALUES
mkdir -p gpurun_out
for v in $VALUES; do
  env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --steps 120 > gpurun_out/bench_env_$v.json 2> gpurun_out/bench_env_$v.err; tail -c 300 gpurun_out/bench_env_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_env_$v.json').read())
r=d['roofline']
print('[$VAR=$v] value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'frac',round(r['frac'],3), {k:round(x,4) for k,x in r['stage_ms'].items()})"
done
