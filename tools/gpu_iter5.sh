#!/usr/bin/env bash
bash tools/gpu_iter4.sh
bash tools/gpu_ll2.sh > gpurun_out/ll.txt 2>&1; python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=(r[4].split('(')[0][:40], r[8], r[7])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1000
for k,(n,t) in agg.items(): print(f"{k[0]:42s} {k[1]:16s} {k[2]:14s} n={n:3d} avg={t/n:7.1f} us")
PY
