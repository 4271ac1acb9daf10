"""Knock-out matrix of conv_plane_kernel on the data-gradient launch shapes of the training benchmark (HERMES-CR-120 grid,
batch 64; K = hi|lo dOut pair rows): CM_PLANE_DBG bits as tools/knockout.py.  usage: python tools/dgrad_knock.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_REPS"] = "20"
shapes = [(0, 64, 8, 28, 24, 64, 32, 0, False), (0, 64, 8, 28, 24, 64, 96, 0, False), (0, 64, 8, 28, 24, 64, 64, 0, False)]
for dbg in (0, 4, 64, 3, 68, 67, 71):
    os.environ["CM_PLANE_DBG"] = str(dbg)
    print(f"--- plane terms=2 dbg={dbg}", file=sys.stderr, flush=True)
    for s in shapes:
        run_conv(nat, *s[:8], 2, s[8], impl=2)
