#!/usr/bin/env bash
# light ncu pass: duration + a few throughput counters for every kernel of one denoiser step
mkdir -p gpurun_out
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__waves_per_multiprocessor \
  --clock-control none -k 'regex:conv_umma|gn_silu|attn_core|first_conv|final_conv' -s 136 -c 67 --csv --log-file gpurun_out/launches.csv \
  python tools/profile_ops.py 64 > gpurun_out/ncu_light.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_light.log; wc -c gpurun_out/launches.csv
