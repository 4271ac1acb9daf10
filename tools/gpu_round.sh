#!/usr/bin/env bash
# round checkpoint: tests, smoke, bench (both arms), launch list, one full-set ncu capture of the top kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "== tests exit $?"; tail -3 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -c 2500 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref exit $?"; tail -c 600 gpurun_out/bench_ref.json
python tools/profile_ops.py 64 gpurun_out/ops.json > gpurun_out/ops.txt 2>&1
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k 'regex:conv_umma|conv_plane|gn_|attn_|first_conv|final_conv' -s 126 -c 63 --csv --log-file gpurun_out/launches.csv python tools/profile_ops.py 64 > gpurun_out/ncu_ll.log 2>&1
echo "launchlist exit $?"
REPS=1 python tools/profile_ops.py 64 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:conv_plane' -s 33 -c 3 -o gpurun_out/conv_plane_full python tools/profile_ops.py 64 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
