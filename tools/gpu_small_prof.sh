mkdir -p gpurun_out
timeout 60 python tools/small_frame_probe.py 5 15 upload 2>&1 | tail -2 || exit 1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_tests_n.log
cat gpurun_out/r02_tests_n.log
MKF_SMALL_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame_small -s 1 -c 1 -f -o gpurun_out/r02_prof_k_frame_small python tools/small_frame_probe.py 262144 15 reset 2>&1 | tail -3
ls -la gpurun_out/r02_prof_k_frame_small.ncu-rep
