mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'stage',d['roofline']['stage_ms'])"
python tools/soak_parity.py 64 500 200 0
python tools/soak_parity.py 4 8192 40 0 slot
python tools/bench_configs.py 4
