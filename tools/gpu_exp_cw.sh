# A/B of the storage chunk width (SlotLay::H): default build vs `make -C mkfbodytracker_pdaf_b200/csrc variant NAME=cw4 DEFS=-DMKF_CW=4` (and NAME=cw8 DEFS=-DMKF_CW=8)
mkdir -p gpurun_out
V=${V:-cw4}
MKF_LIB_VARIANT=$V timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "" $V cw8; do
  MKF_LIB_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_exp_$v.json 2> gpurun_out/bench_exp_$v.err; tail -c 300 gpurun_out/bench_exp_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_exp_$v.json').read())
r=d['roofline']; e=r['every_slot_computed']
print('[$v] value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'frac',round(r['frac'],3), {k:round(x,4) for k,x in r['stage_ms'].items()}, 'every-slot ms',round(e['ms_per_step'],4),'kernel',round(e['kernel_ms'],4),'frac',round(e['frac'],3))"
done
