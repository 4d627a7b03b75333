mkdir -p gpurun_out
python tools/bench_configs.py 2 2lit > gpurun_out/configs_lit_r01.jsonl 2> gpurun_out/configs_r01.err
cat gpurun_out/configs_lit_r01.jsonl; tail -5 gpurun_out/configs_r01.err
