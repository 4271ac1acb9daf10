#!/usr/bin/env bash
# round-2 checkpoint A: parity suite (report mode: every per-tensor gradient error is printed), smoke,
# bench with the new legs, chain-graph kinds, dgrad operand terms, full-resolution weight terms
mkdir -p gpurun_out
CM_TEST_REPORT_ONLY=1 timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/a_tests.log 2>&1; echo "== tests exit $?"; grep -E "passed|failed|error" gpurun_out/a_tests.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/a_smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "== bench exit $?"; tail -c 1500 gpurun_out/a_bench.json
CROWDMOD_CHAIN_GRAPH=step timeout 600 python bench.py --steps 5 --warmup 3 --no-train --no-extras --no-cpu-baseline > gpurun_out/a_bench_stepgraph.json 2> gpurun_out/a_bench_stepgraph.err; echo "== bench(step graph) exit $?"; head -c 400 gpurun_out/a_bench_stepgraph.json
CROWDMOD_DGRAD_TERMS=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/a_bench_dgrad1.json 2> gpurun_out/a_bench_dgrad1.err; echo "== bench(dgrad terms 1) exit $?"
CROWDMOD_FULLRES_TERMS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --no-header -p no:cacheprovider -s -k "golden or oracle" > gpurun_out/a_tests_fullres1.log 2>&1; echo "== tests(fullres terms 1) exit $?"; grep -E "rel-L2|passed|failed" gpurun_out/a_tests_fullres1.log | tail -20
CROWDMOD_FULLRES_TERMS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-train --no-extras --no-cpu-baseline > gpurun_out/a_bench_fullres1.json 2> gpurun_out/a_bench_fullres1.err; echo "== bench(fullres terms 1) exit $?"; head -c 300 gpurun_out/a_bench_fullres1.json
CROWDMOD_WEIGHT_TERMS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --no-header -p no:cacheprovider -s -k "golden or oracle" > gpurun_out/a_tests_terms1.log 2>&1; echo "== tests(weight terms 1) exit $?"; grep -E "rel-L2|passed|failed" gpurun_out/a_tests_terms1.log | tail -20
CROWDMOD_WEIGHT_TERMS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-train --no-extras --no-cpu-baseline > gpurun_out/a_bench_terms1.json 2> gpurun_out/a_bench_terms1.err; echo "== bench(weight terms 1) exit $?"; head -c 300 gpurun_out/a_bench_terms1.json
