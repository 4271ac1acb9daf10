"""Knock-out matrix for conv_plane_kernel on the ATC full-resolution shapes (CM_PLANE_DBG bits:
1 no A loads, 2 no B loads, 4 no MMA, 8 no stores/residual reads, 64 no epilogue)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
for dbg in (0, 1, 2, 3, 4, 8, 64, 68, 67, 71):
    os.environ["CM_PLANE_DBG"] = str(dbg)
    print(f"--- dbg={dbg}", file=sys.stderr, flush=True)
    run_conv(nat, 0, 64, 8, 12, 36, 32, 32, 0, 2, True, impl=2)
    run_conv(nat, 0, 64, 8, 12, 36, 96, 32, 96, 2, False, impl=2)
    run_conv(nat, 0, 64, 4, 6, 18, 64, 64, 0, 2, True, impl=2)
