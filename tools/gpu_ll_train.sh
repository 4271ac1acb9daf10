#!/usr/bin/env bash
# ncu launch list (durations) of one training step (forward + native backward), HERMES-CR-120 shape, batch 64
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 640 -c 330 --csv --log-file gpurun_out/launches_train.csv python tools/bwd_trace.py > gpurun_out/ncu_ll_train.log 2>&1
echo "exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_train.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split('(')[0][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1].replace(',',''))/1000
tot=sum(a[1] for a in agg.values())
print('total us',tot,'launches',len(rows))
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:40]: print(f"{k:62s} n={n:4d} total={t:9.1f} us avg={t/n:7.1f}")
PY
