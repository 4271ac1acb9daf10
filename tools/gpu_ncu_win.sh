#!/usr/bin/env bash
# targeted ncu capture (selected sections, no source) of a window of kernels; raw CSV exported on the box
# $1 = tag, $2 = kernel regex, $3 = skip, $4 = count
mkdir -p gpurun_out
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats \
    --clock-control none -k "regex:$2" -s $3 -c $4 -o gpurun_out/win_$1 python tools/profile_ops.py 64 > gpurun_out/ncu_win_$1.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/win_$1.ncu-rep --page raw --csv > gpurun_out/win_$1_raw.csv 2>/dev/null
ls -la gpurun_out/win_$1*
