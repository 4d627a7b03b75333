# ncu evidence for profiles/ (round 2): launch list of the headline bench command and full captures of the per-step
# kernels of the run-length pipeline in steady state.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
R=${ROUND:-r02}
TAG=${TAG:-a}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --repeats 1 --headline-only --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
echo "launch list rc=$?"
for k in ${KERNELS:-k_frame_heads k_slot_update_heads_direct k_resample_runs k_estimate_runs}; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 20 -c 2 -f -o gpurun_out/prof_${k}_${R}${TAG} \
      python bench.py --steps 24 --warmup 3 --repeats 1 --headline-only --no-cpu-baseline > gpurun_out/b_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
