#!/usr/bin/env bash
# per-op event timings + ncu launch list + ncu full capture of one denoiser step (B=64)
mkdir -p gpurun_out
python tools/profile_ops.py 64 gpurun_out/ops.json > gpurun_out/ops.txt 2>&1 || { tail -5 gpurun_out/ops.txt; exit 1; }
tail -75 gpurun_out/ops.txt
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:conv_umma|gn_silu|attn_core|first_conv|final_conv' -s 136 -c 68 -o gpurun_out/step_full python tools/profile_ops.py 64 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
