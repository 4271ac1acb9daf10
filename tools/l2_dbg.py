import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
split = sys.argv[1] if len(sys.argv) > 1 else "0"
if split == "0": os.environ["CM_NO_SPLITK"] = "1"
for stages in ("8",):
    os.environ["CM_DBG_STAGES"] = stages
    for skip in (0, 7):
        os.environ["CM_DBG_SKIP"] = str(skip)
        print(f"--- split={split} stages={stages} skip={skip}", file=sys.stderr, flush=True)
        run_conv(nat, 0, 64, 2, 3, 9, 128, 128, 0, 2, True, impl=0)
