#!/usr/bin/env bash
for shape in coarse128 full32; do
  for skip in 0 15; do
    echo "=== $shape skip=$skip"
    CM_DBG_TRACE=1 CM_DBG_SKIP=$skip python tools/conv_probe.py $shape 2 2>&1 | grep CM_TRACE | head -40
  done
done
