#!/usr/bin/env bash
# round-2 checkpoint C: precise training forward (hi|lo activation operands) -- parity report + training bench
mkdir -p gpurun_out
CM_TEST_REPORT_ONLY=1 timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/c_tests.log 2>&1; echo "== tests exit $?"; grep -E "passed|failed|error" gpurun_out/c_tests.log | tail -3
grep -E "global grad|^TENSOR" gpurun_out/c_tests.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/c_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "== bench exit $?"; python -c "
import json; b=json.loads(open('gpurun_out/c_bench.json').read().strip().splitlines()[-1]); print(b['value'], {k:v for k,v in b['train'].items() if k not in ('workload',)})"
CROWDMOD_TRAIN_ACT_TERMS=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/c_bench_act1.json 2> gpurun_out/c_bench_act1.err; echo "== bench(act terms 1) exit $?"; python -c "
import json; b=json.loads(open('gpurun_out/c_bench_act1.json').read().strip().splitlines()[-1]); print(b['value'], {k:v for k,v in b['train'].items() if k not in ('workload',)})"
