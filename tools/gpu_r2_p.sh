#!/usr/bin/env bash
# checkpoint: full GPU suite, smoke, bench with the driver's flags
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/p_tests.log 2>&1; echo "== tests exit $?"; tail -3 gpurun_out/p_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/p_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "== bench exit $?"; python - <<'PY'
import json
b = json.loads(open('gpurun_out/p_bench.json').read().strip().splitlines()[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['roofline']['frac'], b['roofline']['per_kernel_ms_per_step'])
print({k: v for k, v in b['train'].items() if k in ('step_ms', 'fwd_ms', 'bwd_ms', 'adam_ms', 'samples_per_s')})
print({k: (v['sequences_per_s'], v['denoiser_step_ms']) for k, v in b['extra'].items()})
PY
