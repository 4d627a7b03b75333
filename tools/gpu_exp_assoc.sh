# association path: parity tests, then config 3 with and without the materialised per-slot columns, then config 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 1 0; do
  MKF_ASSOC_MATERIALISE=$m timeout 600 python tools/bench_configs.py 3 > gpurun_out/cfg3_mat$m.jsonl 2> gpurun_out/cfg3_mat$m.err; tail -c 300 gpurun_out/cfg3_mat$m.err
  python -c "
import json
for l in open('gpurun_out/cfg3_mat$m.jsonl'):
    d=json.loads(l); print('[materialise=$m]', d['config'][:60], 'assoc_only_ms', round(d['assoc_only_ms'],4), 'assoc+update ms', round(d['assoc_plus_update_ms'],4), 'frame-updates/s', round(d['frame_updates_per_s']))"
done
timeout 600 python bench.py --no-cpu-baseline --steps 120 > gpurun_out/bench_after_assoc.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_after_assoc.json').read()); r=d['roofline']
print('config 2: value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'frac',round(r['frac'],3), {k:round(x,4) for k,x in r['stage_ms'].items()})"
