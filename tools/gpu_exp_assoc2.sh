mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/bench_configs.py 3 > gpurun_out/cfg3_new.jsonl 2> gpurun_out/cfg3_new.err; tail -c 300 gpurun_out/cfg3_new.err
python -c "
import json
for l in open('gpurun_out/cfg3_new.jsonl'):
    d=json.loads(l); print(d['config'][:60], 'assoc_only_ms', round(d['assoc_only_ms'],4), 'assoc+update ms', round(d['assoc_plus_update_ms'],4), 'frame-updates/s', round(d['frame_updates_per_s']), 'cand weights/s', round(d['candidate_weights_per_s']))"
