"""Bring-up probe: times cm_op_conv3d on the ATC layer shapes under CM_DBG_* knobs."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {   # name: (mode, B, D, H, W, cin, cout, cin_extra)
    "full32": (0, 64, 8, 12, 36, 32, 32, 0),
    "full64to32": (0, 64, 8, 12, 36, 64, 32, 0),
    "mid64": (0, 64, 4, 6, 18, 64, 64, 0),
    "coarse128": (0, 64, 2, 3, 9, 128, 128, 0),
    "inproj": (3, 64, 2, 3, 9, 128, 384, 0),
    "up64": (2, 64, 4, 6, 18, 64, 64, 0),
}

def child(name, terms):
    import torch
    import crowdmod_ddpm_4d_b200._native as nat
    mode, B, D, H, W, cin, cout, cx = SHAPES[name]
    g = torch.Generator(device="cuda").manual_seed(0)
    act = torch.randn(B, D, H, W, cin, device="cuda", generator=g).half()
    k = 1 if mode == 3 else 3
    w = torch.randn(cout, cin, k, k, k, device="cuda", generator=g)
    bias = torch.randn(cout, device="cuda", generator=g)
    s = 2 if mode == 2 else 1
    out = torch.zeros(B, D * s, H * s, W * s, cout, device="cuda")
    nat.check(nat.lib().cm_op_conv3d(mode, nat.ptr(act), B, D, H, W, cin, None, 0, nat.ptr(w), None, nat.ptr(bias),
                                     cout, terms, None, nat.ptr(out), None, 0, nat.current_stream()))
    torch.cuda.synchronize()

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1], int(sys.argv[2]))
        sys.exit(0)
    for name in os.environ.get("PROBE_SHAPES", ",".join(SHAPES)).split(","):
        envs = [{}, {"CM_DBG_SKIP": "16"}, {"CM_DBG_SKIP": "32"}, {"CM_DBG_SKIP": "48"}, {"CM_DBG_SKIP": "15"},
                {"CM_DBG_SKIP": "31"}, {"CM_DBG_SKIP": "47"}, {"CM_DBG_SKIP": "79"}, {"CM_DBG_SKIP": "95"}, {"CM_DBG_SKIP": "127"}]
        if os.environ.get("PROBE_SET") == "1":
            envs = [{}, {"CM_DBG_STAGES": "2"}, {"CM_DBG_STAGES": "6"}, {"CM_DBG_SKIP": "1"}, {"CM_DBG_SKIP": "2"},
                    {"CM_DBG_SKIP": "3"}, {"CM_DBG_SKIP": "4"}, {"CM_DBG_SKIP": "8"}, {"CM_DBG_SKIP": "15"}, {"TERMS": "1"}]
        for env in envs:
            e = dict(os.environ, CM_DBG_REPS="50", **{k: v for k, v in env.items() if k != "TERMS"})
            terms = env.get("TERMS", "2")
            r = subprocess.run([sys.executable, __file__, name, terms], env=e, capture_output=True, text=True)
            line = [l for l in r.stderr.splitlines() if "CM_DBG" in l]
            print(name, env, line[-1] if line else ("FAILED " + r.stderr[-300:]), flush=True)
