mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:^(k_resample_block|k_estimate|k_share_keys)\$" -s 60 -c 3 -f -o gpurun_out/prof_light \
      python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_light.log 2>&1
ls -la gpurun_out/*.ncu-rep
