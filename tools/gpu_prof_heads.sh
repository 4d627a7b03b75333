mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:^(k_slot_update_heads_direct)\$" -s 60 -c 1 -f -o gpurun_out/prof_heads \
      python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_heads.log 2>&1
ls -la gpurun_out/*.ncu-rep
