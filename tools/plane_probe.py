"""Bring-up probe for the plane-tile conv (conv_plane.cuh): correctness vs torch for both UMMA
descriptor base-offset conventions, then timings of the ATC full-resolution layer shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat  # noqa: E402
from tests.test_gpu_ops import run_conv, rel_l2  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
CASES = [
    (0, 2, 8, 12, 36, 32, 32, 0, True),
    (0, 2, 8, 12, 36, 64, 32, 0, False),
    (0, 1, 8, 12, 36, 96, 32, 0, False),
    (0, 2, 8, 12, 36, 32, 32, 96, False),
    (0, 2, 4, 6, 18, 64, 64, 32, False),
    (0, 1, 4, 6, 18, 192, 64, 0, False),
    (0, 1, 8, 28, 24, 32, 32, 0, True),
]
for mode in ("1", "0"):
    os.environ["CM_PLANE_BASEOFF"] = mode
    for case in CASES:
        m, B, D, H, W, cin, cout, cx, resid = case
        for terms in (2, 1):
            try:
                out32, out16, ref, flag = run_conv(nat, m, B, D, H, W, cin, cout, cx, terms, resid, impl=2)
                print(f"baseoff={mode} case={case} terms={terms}: rel-L2 {rel_l2(out32, ref):.3e} flag={flag}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"baseoff={mode} case={case} terms={terms}: ERROR {e}", flush=True)
if len(sys.argv) > 1:
    os.environ["CM_PLANE_BASEOFF"] = sys.argv[1]
    os.environ["CM_DBG_REPS"] = "20"
    for case in [(0, 64, 8, 12, 36, 32, 32, 0, True), (0, 64, 8, 12, 36, 64, 32, 0, False),
                 (0, 64, 8, 12, 36, 96, 32, 0, False), (0, 64, 8, 12, 36, 32, 32, 96, False),
                 (0, 64, 4, 6, 18, 192, 64, 0, False), (0, 64, 4, 6, 18, 64, 64, 32, False)]:
        m, B, D, H, W, cin, cout, cx, resid = case
        run_conv(nat, m, B, D, H, W, cin, cout, cx, 2, resid, impl=2)
        run_conv(nat, m, B, D, H, W, cin, cout, cx, 2, resid, impl=0)
