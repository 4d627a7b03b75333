# ncu --set full captures of the secondary kernels (association, legacy pf2D, small-N bank kernels)
mkdir -p gpurun_out
R=${ROUND:-r01}
ncu --set full --clock-control none --import-source on -k "regex:^k_pf2d" -s 4 -c 4 -f -o gpurun_out/prof_pf2d_$R \
    python tools/bench_configs.py pf2d > gpurun_out/p2_pf2d.log 2>&1; echo "pf2d rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:^k_assoc" -s 8 -c 4 -f -o gpurun_out/prof_assoc_$R \
    python tools/bench_configs.py 3 > gpurun_out/p2_assoc.log 2>&1; echo "assoc rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:^(k_estimate_small|k_resample_small|k_indicator_bounds|k_slot_update_repair)$" -s 8 -c 8 -f -o gpurun_out/prof_small_$R \
    python tools/bench_configs.py 5 > gpurun_out/p2_small.log 2>&1; echo "small rc=$?"
ls -la gpurun_out/*.ncu-rep
