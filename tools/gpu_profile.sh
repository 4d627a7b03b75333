# ncu evidence for profiles/: launch list of the bench command, full captures of the three per-step kernels,
# and a launch list of the secondary configs.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
R=${ROUND:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
for k in k_slot_update k_resample_block k_estimate; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" -s 3 -c 2 -f -o gpurun_out/prof_${k}_$R \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_configs_$R.csv \
    python tools/bench_configs.py 4 5 2b pf2d 3 > gpurun_out/cfg_ncu.log 2>&1
echo "configs rc=$?"
ls -la gpurun_out | tail -12
