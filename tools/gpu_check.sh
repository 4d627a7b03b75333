# quick validation: GPU tests, smoke, bench, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/tests_chk.log 2>&1; tail -3 gpurun_out/tests_chk.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err; tail -c 300 gpurun_out/bench_chk.err
python -c "import json; d=json.loads(open('gpurun_out/bench_chk.json').read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'stage',d['roofline']['stage_ms'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | cut -c1-300
