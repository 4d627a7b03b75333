mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_chk.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_chk.json').read()); r=d['roofline']
print('config 2: value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'sync',round(d['e2e']['sync_every_step']['value']),'frac',round(r['frac'],3), {k:round(x,4) for k,x in r['stage_ms'].items()})"
