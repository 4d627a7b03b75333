mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err; tail -c 600 gpurun_out/bench_exp.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_exp.json').read())
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['sync_every_step'],'launches',d['gpu_launches'])
r=d['roofline']; print(r['frac'], r['stage_ms']); print(r['every_slot_computed'])"
