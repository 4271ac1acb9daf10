"""conv_res32_kernel on the ATC level-0 shapes: launch time (CM_DBG_REPS), with and without the CTA event trace
(CM_PLANE_TRACE).  usage: python tools/res32_probe.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import crowdmod_ddpm_4d_b200._native as nat
    from tests.test_gpu_ops import run_conv
    os.environ["CM_DBG_REPS"] = "20"
    for s in [(0, 64, 8, 12, 36, 32, 32, 0, True), (0, 64, 8, 12, 36, 32, 32, 96, False), (0, 64, 8, 12, 36, 32, 32, 64, False),
              (0, 1280, 8, 12, 36, 32, 32, 0, True)]:
        run_conv(nat, *s[:8], 2, s[8], impl=3)
else:
    for env in ({}, {"CM_PLANE_TRACE": "1"}):
        print("=== trace:", "on" if env else "off", flush=True)
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        for line in r.stderr.splitlines():
            if line.startswith("CM_DBG") or line.startswith("CTA_TIMES"):
                print("   ", line.replace("CM_DBG res32 ", "")[:170])
        if r.returncode:
            print(r.stderr[-2000:])
