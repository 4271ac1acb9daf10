#!/usr/bin/env bash
# full-set ncu capture with source counters of a few launches: $1 tag, $2 regex, $3 skip, $4 count
mkdir -p gpurun_out
REPS=1 ncu --set full --import-source on --clock-control none --cache-control none -k "regex:$2" -s $3 -c $4 -o gpurun_out/src_$1 python tools/profile_ops.py 64 > gpurun_out/ncu_src_$1.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/src_$1.ncu-rep
