# short-track frame in one launch (k_frame_small): parity tests, then A/B of bank mode / config 5 against the five-launch frame
mkdir -p gpurun_out
timeout 60 python tools/small_frame_probe.py 5 15 upload 2>&1 | tail -2 || exit 1
timeout 300 python -m pytest tests -m gpu -x -q -k "short_track or cholesky or degenerate or config5 or wraps or bank or teacher" 2>&1 | tail -15 > gpurun_out/r02_tests_small.log
cat gpurun_out/r02_tests_small.log
grep -q passed gpurun_out/r02_tests_small.log || exit 1
MKF_SMALL_FUSED=0 timeout 300 python tools/bench_configs.py 2b 5 > gpurun_out/r02_small_ab_off.jsonl 2> gpurun_out/small_ab_off.err
MKF_SMALL_FUSED=1 timeout 300 python tools/bench_configs.py 2b 5 > gpurun_out/r02_small_ab_on.jsonl 2> gpurun_out/small_ab_on.err
tail -c 1500 gpurun_out/small_ab_on.err
cat gpurun_out/r02_small_ab_off.jsonl gpurun_out/r02_small_ab_on.jsonl | cut -c1-360
MKF_SMALL_FUSED=1 timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_frame_small -s 1 -c 1 python tools/small_frame_probe.py 262144 15 reset 2>&1 | grep -E "inst_executed|duration"
