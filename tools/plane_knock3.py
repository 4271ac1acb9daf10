import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
for dbg in (71, 71 + 128, 67, 67 + 128):
    for terms in (2, 1):
        os.environ["CM_PLANE_DBG"] = str(dbg)
        print(f"--- dbg={dbg} terms={terms}", file=sys.stderr, flush=True)
        try:
            run_conv(nat, 0, 64, 8, 12, 36, 32, 32, 0, terms, True, impl=2)
        except Exception as e:
            print("ERR", str(e)[:100], file=sys.stderr)
