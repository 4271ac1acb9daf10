"""Per-op device times of one denoiser step (CUDA events around every launch) at the bench batch.
usage: [WORKLOAD=atc|atc_medium|ethucy|hermes] python tools/profile_ops.py [batch] [out.json]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, synthetic_macroprops  # noqa: E402
from crowdmod_ddpm_4d_b200 import _native as nat  # noqa: E402
from crowdmod_ddpm_4d_b200.models.backbones.unet import UNet  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    out = sys.argv[2] if len(sys.argv) > 2 else None
    reps = int(os.environ.get("REPS", "5"))
    wl = WORKLOADS[os.environ.get("WORKLOAD", "atc")]
    ATC, ROWS, COLS, PAST, FUT = wl["unet"], wl["rows"], wl["cols"], wl["past"], wl["fut"]
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    net = UNet(**ATC).to(dev).eval()
    past = synthetic_macroprops(n, 3, ROWS, COLS, PAST, 1234, dev)
    x = torch.randn(n, 3, ROWS, COLS, FUT, device=dev)
    t = torch.full((n,), 500, device=dev, dtype=torch.long)
    with torch.no_grad():
        net(x, t, past)
    plan = net._plan(ROWS, COLS, PAST, FUT)
    lib = nat.lib()
    nops = lib.cm_unet_op_count(plan.handle)
    eps = torch.empty_like(x)
    ms = (C.c_float * nops)()
    best = None
    for _ in range(reps):
        nat.check(lib.cm_unet_profile_forward(plan.handle, nat.ptr(x), nat.ptr(t), nat.ptr(past), nat.ptr(eps), n,
                                              nat.current_stream(), ms, nops))
        cur = list(ms)
        best = cur if best is None else [min(a, b) for a, b in zip(best, cur)]
    rows = []
    tag = C.create_string_buffer(128)
    ty = C.c_int()
    fl = C.c_double()
    kinds = {0: "first", 1: "gn", 2: "conv", 3: "attn", 4: "final"}
    for i in range(nops):
        lib.cm_unet_op_info(plan.handle, i, tag, 128, C.byref(ty), C.byref(fl))
        gf = fl.value * n / 1e9
        rows.append({"i": i, "tag": tag.value.decode(), "kind": kinds[ty.value], "us": best[i] * 1e3, "gflop": gf,
                     "tflops": gf / best[i] if best[i] > 0 else 0.0})
    for r in rows:
        print(f"{r['i']:3d} {r['kind']:5s} {r['tag']:45s} {r['us']:8.1f} us {r['gflop']:8.2f} GF {r['tflops']:7.1f} TF/s")
    print("total us", sum(r["us"] for r in rows))
    if out:
        json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
