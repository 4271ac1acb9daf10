"""Bring-up: run the same samples through the plan at two batch sizes and report, op by op, how far the
intermediate tensors of the first samples are apart (finds the first op whose large-batch path differs).
  python tools/batch_divergence.py [small=64] [big=1280]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crowdmod_ddpm_4d_b200 import _native as nat  # noqa: E402
from crowdmod_ddpm_4d_b200.models.backbones.unet import UNet  # noqa: E402

KW = dict(input_channels=3, output_channels=3, num_res_blocks=1, base_channels=32, base_channels_multiples=[1, 2, 4],
          apply_attention=[False, False, True, False], dropout_rate=0.1, time_multiple=4, condition="Past")


def dump(net, plan, nsamp):
    lib = nat.lib()
    out = {}
    tag = C.create_string_buffer(128)
    ty = C.c_int()
    fl = C.c_double()
    for i in range(lib.cm_unet_op_count(plan.handle)):
        lib.cm_unet_op_info(plan.handle, i, tag, 128, C.byref(ty), C.byref(fl))
        ti = lib.cm_unet_debug_op_tensor(plan.handle, i)
        if ti < 0:
            continue
        for kind, dt in ((32, torch.float32), (16, torch.float16)):
            buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
            c, p = C.c_int(), C.c_int()
            rc = lib.cm_unet_debug_tensor_read(plan.handle, ti, kind, 0, nsamp, nat.ptr(buf), buf.numel(), C.byref(c),
                                               C.byref(p), nat.current_stream())
            if rc != 0:
                continue
            n = nsamp * p.value * c.value
            out[(i, tag.value.decode(), kind)] = buf.view(dt)[:n].clone().float()
    torch.cuda.synchronize()
    return out


def main():
    small = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    big = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
    torch.manual_seed(42)
    net = UNet(**KW).cuda().eval()
    g = torch.Generator().manual_seed(1280)
    x = torch.randn(small, 3, 12, 36, 3, generator=g).cuda()
    p = torch.randn(small, 3, 12, 36, 5, generator=g).cuda()
    t = torch.randint(0, 1000, (small,), generator=g).cuda()
    reps = big // small
    plan = net._plan(12, 36, 5, 3)
    with torch.no_grad():
        net(x, t, p)
        a = dump(net, plan, 2)
        net(x.repeat(reps, 1, 1, 1, 1), t.repeat(reps), p.repeat(reps, 1, 1, 1, 1))
        b = dump(net, plan, 2)
    for k in a:
        if k not in b:
            continue
        d = (a[k] - b[k]).double().norm() / a[k].double().norm().clamp_min(1e-30)
        print(f"op {k[0]:3d} {k[1]:48s} fp{k[2]}  rel diff {d.item():.3e}")


if __name__ == "__main__":
    main()
