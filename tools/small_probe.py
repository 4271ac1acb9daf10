"""Warm-L2 timings of the small kernels (events around 50 back-to-back launches)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crowdmod_ddpm_4d_b200._native as nat
lib = nat.lib()

def timeit(fn, reps=50):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps

def gn(B, pixels, c0, c1):
    s0 = torch.randn(B, pixels, c0, device="cuda"); s1 = torch.randn(B, pixels, c1, device="cuda") if c1 else None
    C_ = c0 + c1
    g = torch.ones(C_, device="cuda"); b = torch.zeros(C_, device="cuda")
    out = torch.empty(B, pixels, C_, device="cuda", dtype=torch.half)
    st = nat.current_stream()
    f = lambda: nat.check(lib.cm_op_gn_silu(nat.ptr(s0), c0, nat.ptr(s1), c1, nat.ptr(g), nat.ptr(b), B, pixels, 1e-5, 1, nat.ptr(out), None, st))
    print(f"gn B={B} px={pixels} C={c0}+{c1}: {timeit(f):.1f} us  ({(s0.numel()*4+(s1.numel()*4 if c1 else 0)+out.numel()*2)/1e6:.1f} MB)")

def attn(B, S, C_):
    qkv = torch.randn(B, S, 3 * C_, device="cuda"); ctx = torch.empty(B, S, C_, device="cuda", dtype=torch.half)
    st = nat.current_stream()
    f = lambda: nat.check(lib.cm_op_attn_core(nat.ptr(qkv), nat.ptr(ctx), B, S, C_, 4, st))
    print(f"attn B={B} S={S} C={C_}: {timeit(f):.1f} us")

def first(B):
    x = torch.randn(B, 3, 12, 36, 3, device="cuda"); past = torch.randn(B, 3, 12, 36, 5, device="cuda")
    w = torch.randn(32, 3, 3, 3, 3, device="cuda"); b = torch.randn(32, device="cuda"); out = torch.empty(B, 8, 12, 36, 32, device="cuda")
    st = nat.current_stream()
    f = lambda: nat.check(lib.cm_op_first_conv(nat.ptr(x), nat.ptr(past), nat.ptr(w), nat.ptr(b), nat.ptr(out), B, 12, 36, 5, 3, 3, 32, st))
    print(f"first B={B}: {timeit(f):.1f} us")

def final(B):
    act = torch.randn(B, 8, 12, 36, 32, device="cuda").half(); w = torch.randn(3, 32, 3, 3, 3, device="cuda"); b = torch.randn(3, device="cuda")
    eps = torch.empty(B, 3, 12, 36, 3, device="cuda"); st = nat.current_stream()
    f = lambda: nat.check(lib.cm_op_final_conv(nat.ptr(act), nat.ptr(w), nat.ptr(b), nat.ptr(eps), B, 12, 36, 8, 5, 32, 3, st))
    print(f"final B={B}: {timeit(f):.1f} us")

def empty_kernel():
    x = torch.zeros(1, device="cuda")
    print(f"torch tiny kernel (x+=1): {timeit(lambda: x.add_(1)):.1f} us")

empty_kernel()
for args in [(64, 3456, 32, 0), (64, 3456, 64, 32), (64, 432, 64, 0), (64, 432, 128, 64), (64, 54, 128, 0), (64, 54, 128, 128)]:
    gn(*args)
attn(64, 54, 128); first(64); final(64)
