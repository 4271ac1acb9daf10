#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "== tests exit $?"; tail -3 gpurun_out/tests.log
python tools/small_probe.py 2>&1 | grep "gn B"
PROBE_SET=1 PROBE_SHAPES=full32,full64to32,mid64 python tools/conv_probe.py 2>&1 | sed 's/CM_DBG conv //' | cut -c1-42,95-200
