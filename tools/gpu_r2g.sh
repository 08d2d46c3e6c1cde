#!/bin/bash
tag=${1:-r2g}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_wall_c2.log 2>&1; tail -1 $out/${tag}_tp_wall_c2.log
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1; grep -A3 "chain_p last\|gpu ms" $out/${tag}_quick_c2.log | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_offline_launches.csv python tools/bench_offline.py --M 100000 --D 40 --reps 0 > $out/${tag}_ncu_offline.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$out/${tag}_offline_launches.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ix={n:i for i,n in enumerate(h)}
agg={}
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    k=r[ix["Kernel Name"]][:40]; v=float(r[ix["Metric Value"]])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
u=rows[hi+1][ix["Metric Unit"]]
for k,(n,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:25]: print(f"{k:42s} {n:4d} {v:12.1f} {u}")
PY
