# Round-2 ncu evidence for profiles/: launch list of the default bench command (all legs) and full captures of the
# per-frame kernels in steady state.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
echo "launch list rc=$?"
# headline pipeline, steady state (frames 20+)
for k in k_frame_heads k_slot_update_heads_tma k_resample_runs; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 20 -c 2 -f -o gpurun_out/r02_prof_${k} \
      python bench.py --steps 24 --warmup 3 --repeats 1 --headline-only --no-cpu-baseline > /dev/null 2>&1
  echo "$k rc=$?"
done
# the every-slot kernel (config2_every_slot leg: the first k_slot_update launches of the run belong to it)
ncu --set full --clock-control none --import-source on -k "regex:^k_slot_update$" -s 12 -c 2 -f -o gpurun_out/r02_prof_k_slot_update \
    python bench.py --steps 20 --warmup 3 --repeats 1 --legs none --no-cpu-baseline > /dev/null 2>&1
echo "k_slot_update rc=$?"
ls -la gpurun_out/r02_prof_*.ncu-rep
