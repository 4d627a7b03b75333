mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'stage',d['roofline']['stage_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_cfg45.csv \
    python tools/bench_configs.py 5 2b pf2d > gpurun_out/cfg45_ncu.log 2>&1
echo "configs rc=$?"
python tools/bench_configs.py 2b 5 pf2d 2>&1 | cut -c1-330
