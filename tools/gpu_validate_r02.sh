# end-of-round validation on one B200: all GPU tests, the default bench line, the reference arm, and compute-sanitizer over the
# short-track kernels (k_frame_small and the repair kernel's serial tail)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_tests_final.log
cat gpurun_out/r02_tests_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2>> gpurun_out/r02_bench_final.err
echo "reference arm rc=$?"
K='short_track and (301 or 9-16 or 70-9 or 1-12) or cholesky and 15 or degenerate and 15 or unnormalised and 10'
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/sanitize_${tool}_small_r02.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_${tool}_small_r02.log
done
