"""Event trace of CTA 0 of conv_plane_kernel (CM_PLANE_TRACE): producer / MMA issuer / epilogue warp 2 time line (SM clocks).
usage: python tools/plane_trace.py [dbg ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_PLANE_TRACE"] = "1"
for dbg in [int(a) for a in sys.argv[1:]] or [0, 71]:
    os.environ["CM_PLANE_DBG"] = str(dbg)
    print(f"=== 32->32 full-res, B=64, terms=2, dbg={dbg}", file=sys.stderr, flush=True)
    run_conv(nat, 0, 64, 8, 12, 36, 32, 32, 0, 2, True, impl=int(os.environ.get("TRACE_IMPL", "2")))
