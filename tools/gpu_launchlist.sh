#!/usr/bin/env bash
# ncu launch list (durations only, warm caches) of one denoiser step at the bench batch
mkdir -p gpurun_out
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k 'regex:conv_umma|gn_|attn_core|first_conv|final_conv|temb' -s 138 -c 69 --csv --log-file gpurun_out/launches.csv python tools/profile_ops.py 64 > gpurun_out/ncu_ll.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
