#!/usr/bin/env bash
# round-2 checkpoint D: full GPU suite, smoke, bench with the driver's flags, per-op event timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/d_tests.log 2>&1; echo "== tests exit $?"; tail -5 gpurun_out/d_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/d_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "== bench exit $?"; tail -c 3000 gpurun_out/d_bench.json
python tools/profile_ops.py 64 gpurun_out/d_ops.json > gpurun_out/d_ops.txt 2>&1; tail -80 gpurun_out/d_ops.txt
