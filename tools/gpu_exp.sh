mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
MKF_SHARE_SPLIT=$v timeout 300 python bench.py --no-cpu-baseline --steps 400 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split=$v value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'stage',d['roofline']['stage_ms'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_exp.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
python tools/launch_list.py gpurun_out/launches_exp.csv 2>&1 | grep "k_"
