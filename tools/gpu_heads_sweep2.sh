mkdir -p gpurun_out
out=gpurun_out/r02_heads_sweep2.jsonl
: > $out
run() { label=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-baseline $EXTRA 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(json.dumps({'label':'$label','tracks':d['config']['tracks_per_gpu'],'ms_per_step':d['ms_per_step'],'kernel_ms':r['kernel_ms'],'units':r['units_per_launch'],'frac':r['frac']}))" >> $out
}
EXTRA="--tracks 32768" run "32768 tracks, 64 CTAs/SM" MKF_HEADS_CTAS_PER_SM=64
EXTRA="--tracks 32768" run "32768 tracks, 32 CTAs/SM" MKF_HEADS_CTAS_PER_SM=32
EXTRA="--tracks 16384" run "16384 tracks, 48 CTAs/SM" MKF_HEADS_CTAS_PER_SM=48
EXTRA="--tracks 16384" run "16384 tracks, 24 CTAs/SM" MKF_HEADS_CTAS_PER_SM=24
cat $out
