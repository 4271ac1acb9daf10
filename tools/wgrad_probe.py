"""Weight-gradient kernels on the level-0 layer shapes of the training benchmark (HERMES-CR-120, batch 64) and of ATC:
wgrad_umma_kernel (impl 0) against the plane / halo scheme (impl 2), CM_DBG_REPS launches each.
usage: python tools/wgrad_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["CM_DBG_REPS"] = "20"
import torch
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import _bwd_inputs
for case in [(0, 64, 8, 28, 24, 32, 32, 0), (0, 64, 8, 28, 24, 96, 32, 0), (0, 64, 8, 28, 24, 64, 32, 0), (0, 64, 8, 28, 24, 32, 32, 96),
             (0, 64, 8, 12, 36, 32, 32, 0)]:
    mode, B, D, H, W, cin, cout, cx = case
    act, w, extra, wx, dout = _bwd_inputs(*case)
    for impl in (0, 2):
        dw = torch.empty_like(w)
        dwx = torch.empty_like(wx) if cx else None
        nat.check(nat.lib().cm_op_conv3d_wgrad(mode, nat.ptr(act), B, D, H, W, cin, nat.ptr(extra), cx, nat.ptr(dout), cout,
                                               nat.ptr(dw), nat.ptr(dwx), impl, nat.current_stream()))
        torch.cuda.synchronize()
