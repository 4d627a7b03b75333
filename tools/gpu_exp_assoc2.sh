mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0 1; do
MKF_POSE_CACHE=$c timeout 600 python tools/bench_configs.py 3 > gpurun_out/cfg3_new$c.jsonl 2> gpurun_out/cfg3_new$c.err; tail -c 300 gpurun_out/cfg3_new$c.err
python -c "
import json
for l in open('gpurun_out/cfg3_new$c.jsonl'):
    d=json.loads(l); print('[cache=$c]', d['config'][:58], 'assoc_only', round(d['assoc_only_ms'],4), 'assoc+update', round(d['assoc_plus_update_ms'],4), '+estimate', round(d['assoc_update_estimate_ms'],4), 'frame-updates/s', round(d['frame_updates_per_s']))"
done
timeout 600 python bench.py --no-cpu-baseline --steps 120 > gpurun_out/bench_after_assoc.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_after_assoc.json').read()); r=d['roofline']
print('config 2: value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'frac',round(r['frac'],3), {k:round(x,4) for k,x in r['stage_ms'].items()})"
