# ncu full capture of the literal-alias chain walker in steady state (config 2, 4096 x 500)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:^k_slot_update_chain_dyn" -s 14 -c 1 -f -o gpurun_out/prof_chain_dyn_${TAG:-r02} \
    python bench.py --steps 20 --warmup 3 --repeats 1 --legs none --no-cpu-baseline > gpurun_out/b_ncu_alias.log 2>&1
echo "rc=$?"
