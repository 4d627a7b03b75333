# A/B runs of the heads kernel's launch shape and the batch-size dependence of its roofline fraction (config 2 model)
mkdir -p gpurun_out
out=gpurun_out/r02_heads_sweep.jsonl
: > $out
run() { # label, env..., -- args
  label=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-baseline $EXTRA 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(json.dumps({'label':'$label','tracks':d['config']['tracks_per_gpu'],'ms_per_step':d['ms_per_step'],'value':d['value'],'kernel_ms':r['kernel_ms'],'units':r['units_per_launch'],'frac':r['frac'],'stage_ms':d['stage_ms'],'clocks':d['clocks']}))" >> $out
}
EXTRA="" run "default (128 threads, 12 CTAs/SM)" MKF_X=0
EXTRA="" run "64-thread CTAs" MKF_HEADS_BLOCK=64
EXTRA="" run "64-thread CTAs, 24 per SM" MKF_HEADS_BLOCK=64 MKF_HEADS_CTAS_PER_SM=24
EXTRA="" run "128 threads, 6 CTAs/SM" MKF_HEADS_CTAS_PER_SM=6
EXTRA="--tracks 8192" run "8192 tracks" MKF_X=0
EXTRA="--tracks 16384" run "16384 tracks" MKF_X=0
EXTRA="--tracks 32768" run "32768 tracks" MKF_X=0
EXTRA="--tracks 2048" run "2048 tracks" MKF_X=0
EXTRA="--tracks 32768" run "32768 tracks, 64 CTAs/SM" MKF_HEADS_CTAS_PER_SM=64
EXTRA="--tracks 32768" run "32768 tracks, 32 CTAs/SM" MKF_HEADS_CTAS_PER_SM=32
EXTRA="--tracks 16384" run "16384 tracks, 48 CTAs/SM" MKF_HEADS_CTAS_PER_SM=48
EXTRA="--tracks 16384" run "16384 tracks, 24 CTAs/SM" MKF_HEADS_CTAS_PER_SM=24
cat $out
