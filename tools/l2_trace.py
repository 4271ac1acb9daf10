import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
pass
os.environ["CM_DBG_TRACE"] = "1"
for skip in (0, 7):
    os.environ["CM_DBG_SKIP"] = str(skip)
    print(f"--- skip={skip}", file=sys.stderr, flush=True)
    run_conv(nat, 0, 64, 2, 3, 9, 128, 128, 0, 2, True, impl=0)
