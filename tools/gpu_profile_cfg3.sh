# ncu launch list of the config-3 leg (association + both arm updates at 16 384 persons x 500 slots x 17 candidates)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg3_${TAG:-r02}.csv \
    python bench.py --steps 4 --warmup 3 --repeats 1 --legs 3 --no-cpu-baseline > gpurun_out/b_ncu_cfg3.log 2>&1
echo "rc=$?"
