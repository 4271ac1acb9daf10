#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --no-header -p no:cacheprovider -x -k conv > gpurun_out/tests.log 2>&1
echo "== conv tests exit $?"; tail -3 gpurun_out/tests.log
PROBE_SET=1 python tools/conv_probe.py 2>&1 | grep -E "\{\}|STAGES|'15'|TERMS" | sed 's/CM_DBG conv //' | cut -c1-40,95-200
bash tools/conv_trace.sh 2>&1 | grep -A 8 "coarse128 skip=0" | head -10
