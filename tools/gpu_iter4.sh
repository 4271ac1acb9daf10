#!/usr/bin/env bash
# iteration loop: parity + train tests, per-op timing, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider -x -s > gpurun_out/tests.log 2>&1
echo "== tests exit $?"; grep -v "^$" gpurun_out/tests.log | tail -14 | cut -c1-300
python tools/profile_ops.py 64 gpurun_out/ops.json > gpurun_out/ops.txt 2>&1; tail -70 gpurun_out/ops.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "== bench exit $?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('seq/s',d['value'],'e2e',d['e2e']['value'],'step_ms',d['denoiser_step_ms'],'roof',d['roofline']['frac'],d['roofline']['per_kernel_ms_per_step'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
