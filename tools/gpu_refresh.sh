# refresh of the judged artefacts after a kernel change (a trimmed tools/gpu_round1.sh, ~12 min of box time):
# GPU tests, smoke, bench (+ reference arm, + sharing off), launch list, one ncu --set full pass over the per-frame
# kernels, short soak.  Outputs under gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_$R.log 2>&1; tail -3 gpurun_out/tests_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -c 400 gpurun_out/bench_$R.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2>/dev/null
MKF_DEDUP=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nosharing_$R.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
# frame 40 onwards (steady state): 6 launches per frame, 2 frames
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:^(k_indicator_bounds|k_share_keys|k_slot_update_heads_direct|k_slot_update_repair|k_resample_block|k_estimate)\$" \
    -s 240 -c 12 -f -o gpurun_out/prof_frame_$R \
    python bench.py --steps 44 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_frame.log 2>&1
MKF_DEDUP=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^k_slot_update\$" -s 3 -c 2 -f \
    -o gpurun_out/prof_k_slot_update_$R python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_k_slot_update.log 2>&1
timeout 300 python tools/soak_parity.py 64 500 300 0 > gpurun_out/soak_$R.jsonl 2>&1
timeout 600 python tools/bench_configs.py > gpurun_out/configs_$R.jsonl 2> gpurun_out/configs_$R.err; tail -c 300 gpurun_out/configs_$R.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_$R.json').read())
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
r=d['roofline']; print(r['frac'], r['stage_ms']); print(r['every_slot_computed']); print(d['cpu_baseline'])"
ls -la gpurun_out | tail -20
