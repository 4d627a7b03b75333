# final validation of a round: GPU tests, smoke, bench (+ reference arm), secondary configs, soak, launch list,
# ncu full captures.  Outputs under gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
R=${ROUND:-r01}
python -m pytest tests -m gpu -q > gpurun_out/tests_$R.log 2>&1; tail -3 gpurun_out/tests_$R.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -c 400 gpurun_out/bench_$R.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2>/dev/null
MKF_DEDUP=0 python bench.py --no-cpu-baseline > gpurun_out/bench_nosharing_$R.json 2>/dev/null
MKF_SHARE_SPLIT=0 python bench.py --no-cpu-baseline > gpurun_out/bench_single_launch_sharing_$R.json 2>/dev/null
python tools/bench_configs.py > gpurun_out/configs_$R.jsonl 2> gpurun_out/configs_$R.err
python tools/soak_parity.py 64 500 300 0 > gpurun_out/soak_$R.jsonl 2>&1
python tools/soak_parity.py 16 500 300 1 slot >> gpurun_out/soak_$R.jsonl 2>&1
python tools/soak_parity.py 2048 15 200 0 >> gpurun_out/soak_$R.jsonl 2>&1
python tools/soak_parity.py 4 8192 60 0 slot >> gpurun_out/soak_$R.jsonl 2>&1
python tools/bench_node.py ref 40 > gpurun_out/node_$R.jsonl 2>/dev/null; python tools/bench_node.py dropin 200 >> gpurun_out/node_$R.jsonl 2>/dev/null
python tools/bench_call_latency.py >> gpurun_out/node_$R.jsonl 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
for k in k_slot_update_heads_direct k_share_keys k_resample_block k_estimate; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" -s 40 -c 2 -f -o gpurun_out/prof_${k}_$R \
      python bench.py --steps 44 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$k.log 2>&1
done
MKF_SHARE_SPLIT=0 ncu --set full --clock-control none --import-source on -k "regex:^k_slot_update_shared\$" -s 40 -c 2 -f -o gpurun_out/prof_k_slot_update_shared_$R \
      python bench.py --steps 44 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_k_slot_update_shared.log 2>&1
MKF_DEDUP=0 ncu --set full --clock-control none --import-source on -k "regex:^k_slot_update\$" -s 3 -c 2 -f -o gpurun_out/prof_k_slot_update_$R \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_k_slot_update.log 2>&1
ls gpurun_out | tail -30
