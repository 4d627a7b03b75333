mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r01.json').read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'],'clocks',d['clocks'],'roofline',d['roofline']['frac'], d['cpu_baseline'])"
tail -3 gpurun_out/bench_r01.err
