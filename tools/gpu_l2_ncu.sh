# does the L2-resident policy change what the heads kernel reads from DRAM?  single-pass metrics, caches left alone
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
run() { # tag tracks l2
  MKF_L2_TRACKS=$3 ncu --cache-control none --clock-control none --metrics $M -k regex:heads_direct -s 30 -c 4 --csv \
    --log-file gpurun_out/r02_l2_ncu_$1.csv python bench.py --steps 40 --warmup 5 --headline-only --no-cpu-baseline --tracks $2 > /dev/null 2>&1
  echo "== $1 (tracks $2, MKF_L2_TRACKS=$3)"; grep -v "^==" gpurun_out/r02_l2_ncu_$1.csv | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>5]
h=rows[0]; iN=h.index('Metric Name'); iV=h.index('Metric Value'); iI=h.index('ID')
d={}
for r in rows[1:]: d.setdefault(r[iI],{})[r[iN]]=r[iV]
for k,v in d.items(): print(k, v)
"
}
run t1024_off 1024 0
run t1024_on 1024 1024
run t4096_off 4096 0
run t4096_1200 4096 1200
run t4096_all 4096 4096
