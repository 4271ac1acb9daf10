#!/usr/bin/env bash
# targeted full-set ncu capture of a few kernels (small report): $1 = kernel regex, $2 = skip, $3 = count
mkdir -p gpurun_out
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none -k "regex:$1" -s $2 -c $3 -o gpurun_out/sel_$4 python tools/profile_ops.py 64 > gpurun_out/ncu_sel.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/sel_$4.ncu-rep
