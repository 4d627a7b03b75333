set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/tests_r01.log
tail -15 gpurun_out/tests_r01.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 3000 gpurun_out/bench_r01.json; tail -5 gpurun_out/bench_r01.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
ncu --set full --clock-control none --import-source on -k regex:k_slot_update -s 3 -c 2 -o gpurun_out/prof_slot_r01 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
tail -3 gpurun_out/b_ncu2.log
ls -la gpurun_out
