#!/usr/bin/env bash
# one-shot GPU probe used during bring-up: runs each op-test group in its own process so a
# sticky CUDA error in one group does not hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for k in gn_silu attn_core first_conv final_conv conv_umma_vs_scalar conv_umma_vs_torch; do
  timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$k" -x --no-header -p no:cacheprovider > gpurun_out/ops_$k.log 2>&1
  echo "== $k exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/ops_$k.log
done
