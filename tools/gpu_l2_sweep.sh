# A/B runs of the heads kernel's L2-resident track count (MKF_L2_TRACKS; config 2 model, 4096 x 500 unless --tracks)
mkdir -p gpurun_out
out=gpurun_out/r02_l2_sweep.jsonl
: > $out
run() { # label, env..., EXTRA args
  label=$1; shift
  env "$@" python bench.py --steps 100 --warmup 5 --headline-only --no-cpu-baseline $EXTRA 2>gpurun_out/l2_sweep.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(json.dumps({'label':'$label','tracks':d['config']['tracks_per_gpu'],'ms_per_step':d['ms_per_step'],'value':d['value'],'kernel_ms':r['kernel_ms'],'units':r['units_per_launch'],'frac':r['frac'],'stage_ms':d['stage_ms'],'clocks':d['clocks']}))" >> $out
}
for n in ${L2_LIST:-0 400 800 1200 1600 2000 2800 4096}; do
  EXTRA="" run "MKF_L2_TRACKS=$n" MKF_L2_TRACKS=$n
done
EXTRA="--tracks 1024" run "1024 tracks, hints off" MKF_L2_TRACKS=0
EXTRA="--tracks 1024" run "1024 tracks, all resident" MKF_L2_TRACKS=1024
EXTRA="--tracks 2048" run "2048 tracks, hints off" MKF_L2_TRACKS=0
EXTRA="--tracks 2048" run "2048 tracks, 1200 resident" MKF_L2_TRACKS=1200
cat $out
