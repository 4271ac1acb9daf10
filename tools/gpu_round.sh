#!/usr/bin/env bash
# The round's evidence in one go (run under gpurun from the repo root; everything lands in gpurun_out/, then
# `python tools/refresh_profiles.py r2` files it under profiles/):
#   stage 1  pytest -m gpu, smoke(), bench.py both arms with the driver's flags, per-op event timings
#   stage 2  ncu launch list of the whole profile_ops.py program (durations + DRAM bytes; the last denoiser step is cut out
#            by refresh_profiles.py), only after the same command exited 0 without ncu
#   stage 3  ncu --set full captures (source counters on) of one launch of each top conv instantiation, raw pages
#            exported on the box
# usage: bash tools/gpu_round.sh [stages, default "1 2 3"]
STAGES="${*:-1 2 3}"
ONLY="${CAPS_ONLY:-}"     # optional: space-separated capture tags of stage 3
mkdir -p gpurun_out
if [[ " $STAGES " == *" 1 "* ]]; then
  timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "== tests exit $?"; tail -3 gpurun_out/tests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/smoke.log
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref exit $?"; tail -c 400 gpurun_out/bench_ref.json
  timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; head -c 400 gpurun_out/bench.json; echo
  timeout 300 python tools/profile_ops.py 64 gpurun_out/ops.json > gpurun_out/ops.txt 2>&1; tail -1 gpurun_out/ops.txt
fi
if [[ " $STAGES " == *" 2 "* ]]; then
  REPS=1 timeout 300 python tools/profile_ops.py 64 > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -c 600 --csv --log-file gpurun_out/launches.csv env REPS=1 python tools/profile_ops.py 64 > gpurun_out/ncu_ll.log 2>&1
  echo "== launch list exit $?"
fi
if [[ " $STAGES " == *" 3 "* ]]; then
  # tag | demangled-name regex | launches of that instantiation to skip (the first forward of profile_ops.py is the warm-up)
  while read -r tag regex skip; do
    [[ -z "$tag" ]] && continue
    [[ -n "$ONLY" && " $ONLY " != *" $tag "* ]] && continue
    REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name-base demangled \
        -k "regex:$regex" -s "$skip" -c 1 -f -o "gpurun_out/full_$tag" python tools/profile_ops.py 64 > "gpurun_out/ncu_full_$tag.log" 2>&1
    echo "== ncu full $tag exit $?"
    ncu -i "gpurun_out/full_$tag.ncu-rep" --page raw --csv > "gpurun_out/full_${tag}_raw.csv" 2>/dev/null
  done <<'CAPS'
res32_enc0conv1      conv_res32_kernel                          6
plane_dec6conv1      conv_plane_kernel<.int.32,.*int.2>         2
plane_dec3conv1      conv_plane_kernel<.int.64,.*int.2>         8
umma_coarse          conv_umma_kernel<.int.128,.*int.2>         17
umma_upsample        conv_umma_kernel<.int.64,.*int.2>          3
CAPS
  ls -la gpurun_out/*.ncu-rep
fi
