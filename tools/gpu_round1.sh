mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'stage',d['roofline']['stage_ms'])"
