set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/tests_r01.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 1200 gpurun_out/bench_r01.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r01.json; cut -c1-300 gpurun_out/bench_ref_r01.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_slot_update<|k_resample_block|k_estimate<" -s 12 -c 6 -o gpurun_out/prof_r01 -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
tail -2 gpurun_out/b_ncu2.log
ls -la gpurun_out | head -20
