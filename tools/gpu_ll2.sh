#!/usr/bin/env bash
mkdir -p gpurun_out
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k 'regex:conv_umma|gn_|attn_core|first_conv|final_conv' -s 140 -c 96 --csv --log-file gpurun_out/launches.csv python tools/profile_ops.py 64 > gpurun_out/ncu_ll.log 2>&1
echo "launchlist exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows:
    print(f"{r[0]:>4} {r[4].split('(')[0][:40]:42s} {r[8]:16s} {r[7]:14s} {float(r[-1])/1000:8.1f}")
PY
