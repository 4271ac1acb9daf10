"""Knock-out matrix for the conv kernels on the ATC layer shapes (one parametrised tool; replaces plane_knock*.py).

usage: python tools/knockout.py [plane|umma] [batch]
CM_PLANE_DBG / CM_DBG_SKIP bits: 1 no A loads, 2 no B loads, 4 no MMA, 8 no stores/residual reads, 64 no epilogue.
Timings are printed by the library (CM_DBG_REPS) on stderr."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
kind = sys.argv[1] if len(sys.argv) > 1 else "plane"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.environ["CM_DBG_REPS"] = "20"
if kind == "plane":
    shapes = [(0, B, 8, 12, 36, 32, 32, 0, True), (0, B, 8, 12, 36, 96, 32, 0, False), (0, B, 8, 12, 36, 64, 32, 0, False),
              (0, B, 4, 6, 18, 64, 64, 0, True), (0, B, 4, 6, 18, 192, 64, 0, False)]
    for terms in (2, 1):
        for dbg in (0, 1, 2, 3, 4, 64, 68, 67, 71):
            os.environ["CM_PLANE_DBG"] = str(dbg)
            print(f"--- plane terms={terms} dbg={dbg}", file=sys.stderr, flush=True)
            for s in shapes:
                run_conv(nat, *s[:8], terms, s[8], impl=2)
else:
    shapes = [(0, B, 2, 3, 9, 128, 128, 0, True), (0, B, 2, 3, 9, 256, 128, 0, False), (1, B, 4, 6, 18, 64, 64, 0, False),
              (2, B, 2, 3, 9, 128, 128, 0, False), (2, B, 4, 6, 18, 64, 64, 0, False)]
    for terms in (2, 1):
        for dbg in (0, 1, 2, 3, 4):
            os.environ["CM_DBG_SKIP"] = str(dbg)
            print(f"--- umma terms={terms} dbg={dbg}", file=sys.stderr, flush=True)
            for s in shapes:
                run_conv(nat, *s[:8], terms, s[8], impl=0)
