#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/parity.log 2>&1
echo "== parity exit $?" | tee gpurun_out/summary2.txt; tail -25 gpurun_out/parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/summary2.txt; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "== bench exit $?" | tee -a gpurun_out/summary2.txt; tail -5 gpurun_out/bench1.log
CROWDMOD_WEIGHT_TERMS=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench1_terms1.log 2>&1; echo "== bench terms1 exit $?" | tee -a gpurun_out/summary2.txt; tail -2 gpurun_out/bench1_terms1.log
