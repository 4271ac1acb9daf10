#!/usr/bin/env bash
# backward bring-up: dgrad / wgrad op tests, then the training parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -p no:cacheprovider -k "dgrad or wgrad" > gpurun_out/bwd_ops.log 2>&1
echo "== bwd ops exit $?"; tail -40 gpurun_out/bwd_ops.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/train.log 2>&1
echo "== train exit $?"; tail -60 gpurun_out/train.log | cut -c1-600
