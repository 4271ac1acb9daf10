#!/usr/bin/env bash
# iteration loop + full-set ncu capture of the non-plane kernels of one step (GN, final, conv_umma, attention)
bash tools/gpu_iter4.sh
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:gn_|final_conv|first_conv|conv_umma|attn_' -s 128 -c 64 -o gpurun_out/rest_full python tools/profile_ops.py 64 > gpurun_out/ncu_rest.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_rest.log; ls -la gpurun_out/*.ncu-rep
