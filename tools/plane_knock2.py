"""Stage-count sweep of the plane conv pipeline skeleton (see plane_knock.py for the dbg bits)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
for stages in (2, 3, 4, 6):
    os.environ["CM_PLANE_STAGES"] = str(stages)
    for dbg in (0, 71, 67, 68):
        os.environ["CM_PLANE_DBG"] = str(dbg)
        print(f"--- stages={stages} dbg={dbg}", file=sys.stderr, flush=True)
        run_conv(nat, 0, 64, 8, 12, 36, 32, 32, 0, 2, True, impl=2)
