mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assoc_pf2d.py tests/test_gpu_full_size.py tests/test_gpu_dropin_node.py tests/test_gpu_reference_node.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/bench_configs.py 3 > gpurun_out/cfg3_w.jsonl 2> gpurun_out/cfg3_w.err; tail -c 300 gpurun_out/cfg3_w.err
python -c "
import json
for l in open('gpurun_out/cfg3_w.jsonl'):
    d=json.loads(l); print(d['config'][:58], 'assoc_only', round(d['assoc_only_ms'],4), 'assoc+update', round(d['assoc_plus_update_ms'],4), '+estimate', round(d['assoc_update_estimate_ms'],4), 'frame-updates/s', round(d['frame_updates_per_s']))"
