"""Per-kernel totals of one training step's ncu launch list, split into forward and backward at the first backward-only kernel.
usage: PROFILE_API=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv \
           python tools/bwd_trace.py; python tools/ll_split.py L.csv"""
import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
# split forward / backward at the first backward-only kernel
names=[r[4].split('(')[0] for r in rows]
cut=next((i for i,n in enumerate(names) if 'final_conv_dact' in n or 'scale_from_max' in n), len(rows))
for label,part in (('forward',rows[:cut]),('backward',rows[cut:])):
    agg=collections.OrderedDict()
    for r in part:
        k=r[4].split('(')[0][:60]
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1].replace(',',''))/1000
    tot=sum(a[1] for a in agg.values())
    print(label,'total us',round(tot,1),'launches',len(part))
    for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"  {k:60s} n={n:4d} total={t:9.1f} us avg={t/n:7.1f}")
