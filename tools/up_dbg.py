import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
for ppc in ("1", "2", "4", "8"):
    for stages in ("2", "3", "4"):
        os.environ["CM_PPC"] = ppc
        os.environ["CM_DBG_STAGES"] = stages
        print(f"--- ppc={ppc} stages={stages}", file=sys.stderr, flush=True)
        run_conv(nat, 2, 64, 4, 6, 18, 64, 64, 0, 2, False, impl=0)
