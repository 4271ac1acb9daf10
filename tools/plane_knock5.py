"""Does the number of pipeline stages per unit matter at equal MMAs and bytes?  64->32 layer with BK = 64 vs BK = 32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
for bk32 in (0, 1):
    if bk32: os.environ["CM_PLANE_BK32"] = "1"
    for dbg in (0, 67, 71, 64):
        os.environ["CM_PLANE_DBG"] = str(dbg)
        print(f"--- bk32={bk32} dbg={dbg}", file=sys.stderr, flush=True)
        run_conv(nat, 0, 64, 8, 12, 36, 64, 32, 0, 2, False, impl=2)
        run_conv(nat, 0, 64, 4, 6, 18, 64, 64, 0, 2, True, impl=2)
