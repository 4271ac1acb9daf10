#!/usr/bin/env bash
# launch list of one step: duration + DRAM bytes per launch, warm caches (--cache-control none)
mkdir -p gpurun_out
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k 'regex:conv_umma|conv_plane|gn_|attn_|first_conv|final_conv|pack_first|temb' -s 160 -c 80 --csv --log-file gpurun_out/launches3.csv python tools/profile_ops.py 64 > gpurun_out/ncu_ll3.log 2>&1
echo "launchlist exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches3.csv')) if len(r)>10 and r[0].isdigit()]
# rows: one per (launch, metric)
by=collections.OrderedDict()
for r in rows:
    k=int(r[0]); d=by.setdefault(k,{'name':r[4].split('(')[0][:36],'grid':r[8],'blk':r[7]})
    d[r[-3]]=float(r[-1].replace(',',''))
    d['u_'+r[-3]]=r[-2]
for k,d in by.items():
    print(f"{k:4d} {d['name']:38s} {d['grid']:14s} {d.get('gpu__time_duration.sum',0):9.1f} {d.get('u_gpu__time_duration.sum','')} rd {d.get('dram__bytes_read.sum',0):9.2f} {d.get('u_dram__bytes_read.sum','')} wr {d.get('dram__bytes_write.sum',0):9.2f} {d.get('u_dram__bytes_write.sum','')}")
PY
