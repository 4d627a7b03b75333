mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg3.csv python tools/bench_configs.py 3 > gpurun_out/ll_cfg3.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_cfg3.csv') if l.startswith('"')))
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
seq=[(r[ki].split("(")[0].replace("void ","").strip(), float(r[vi].replace(",",""))) for r in rows[1:]]
# print the steady sequence of the first config (first 200 launches): aggregated by name
agg=collections.OrderedDict()
for n,v in seq[:400]:
    agg.setdefault(n,[]).append(v)
for n,v in agg.items(): print("%-50s %5d  median %9.1f ns" % (n[:50], len(v), sorted(v)[len(v)//2]))
print("sequence sample:", [ (n[:28], round(v/1000,1)) for n,v in seq[150:185]])
PY
