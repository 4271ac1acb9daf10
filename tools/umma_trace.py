"""Time line of one CTA of conv_umma_kernel (CM_DBG_TRACE, %globaltimer ns): coarse-level split-K conv and an UpSample conv."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crowdmod_ddpm_4d_b200._native as nat
from tests.test_gpu_ops import run_conv
os.environ["CM_DBG_TRACE"] = "1"
os.environ["CM_DBG_REPS"] = "20"
for name, s in [("coarse 128->128", (0, 64, 2, 3, 9, 128, 128, 0, True)), ("coarse 256->128", (0, 64, 2, 3, 9, 256, 128, 0, False)),
                ("downsample 64->64", (1, 64, 4, 6, 18, 64, 64, 0, False)), ("upsample 128", (2, 64, 2, 3, 9, 128, 128, 0, False))]:
    print(f"=== {name}", file=sys.stderr, flush=True)
    run_conv(nat, *s[:8], 2, s[8], impl=0)
