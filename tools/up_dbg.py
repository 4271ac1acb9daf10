import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
for skip in (0, 3, 4, 7, 8, 15):
    os.environ["CM_DBG_SKIP"] = str(skip)
    print(f"--- skip={skip}", file=sys.stderr, flush=True)
    run_conv(nat, 2, 64, 4, 6, 18, 64, 64, 0, 2, False, impl=0)
