# A/B of the heads kernel variants on the headline config (4096 x 500 unless --tracks)
mkdir -p gpurun_out
out=gpurun_out/r02_tma_ab.jsonl
: > $out
run() { # label, env..., EXTRA args
  label=$1; shift
  env "$@" python bench.py --steps ${STEPS:-100} --warmup 5 --headline-only --no-cpu-baseline $EXTRA 2>gpurun_out/tma_ab.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['stage_ms']; print(json.dumps({'label':'$label','tracks':d['config']['tracks_per_gpu'],'ms_per_step':round(d['ms_per_step'],5),'value':round(d['value']),'e2e':round(d['e2e']['value']),'kernel_ms':round(r['kernel_ms'],5),'span':round(s.get('slot_kernel_device_span',0),5),'units':r['units_per_launch'],'frac':round(r['frac'],3),'heads':round(s['share_keys'],4),'resample':round(s['normalise_resample'],4)}))" >> $out
}
EXTRA="" run "tile layout, k_slot_update_heads_direct" MKF_HEADS_TMA=0
for c in ${CFGS:-630 440 442 533 543 632 633}; do
EXTRA="" run "TMA cfg $c" MKF_HEADS_TMA_CFG=$c
done
for c in ${NOXS:-}; do
EXTRA="" run "TMA cfg $c, no xs" MKF_HEADS_TMA_CFG=$c MKF_NO_XS=1
done
cat $out
