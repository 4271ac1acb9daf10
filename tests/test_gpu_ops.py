"""Kernel-level parity (GPU): every hand-written kernel of the hot path, called through the
C ABI (cm_op_*), against the PyTorch fp32 op it replaces (TF32 disabled).

Tolerances: the tcgen05 conv takes fp16 operands; with `terms=2` (hi+lo split weights) the only
rounding left is fp32 accumulation order (rel-L2 <= 2e-5 against F.conv3d on the same
fp16-rounded activations); with `terms=1` weight rounding adds ~3e-4 (bound 1e-3).  Against the
scalar restatement of the same packed operands (`impl=1`) it must agree to 1e-5.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import crowdmod_ddpm_4d_b200._native as n
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n.lib()
    return n


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def describe_mismatch(got, ref, tile=128):
    """Diagnostics that localise a wrong tile / column of the implicit GEMM."""
    g = got.reshape(-1, got.shape[-1]).double()
    r = ref.reshape(-1, ref.shape[-1]).double()
    err = (g - r).abs()
    rows = err.amax(dim=1)
    cols = err.amax(dim=0)
    nt = (rows.numel() + tile - 1) // tile
    tiles = [rows[i * tile:(i + 1) * tile].max().item() for i in range(min(nt, 12))]
    return (f"max|err|={err.max().item():.3e} ref_rms={r.pow(2).mean().sqrt().item():.3e} "
            f"first tiles max err={['%.2e' % t for t in tiles]} "
            f"col err (first 16)={['%.1e' % c for c in cols[:16].tolist()]} "
            f"got[0,:4]={g[0,:4].tolist()} ref[0,:4]={r[0,:4].tolist()}")


def torch_conv(mode, act16, extra16, w, wx, bias, resid):
    x = act16.float().permute(0, 4, 1, 2, 3).contiguous()
    if mode == 0:
        y = F.conv3d(x, w, None, stride=1, padding=1)
    elif mode == 1:
        y = F.conv3d(x, w, None, stride=2, padding=1)
    elif mode == 2:
        y = F.conv3d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, stride=1, padding=1)
    else:
        y = F.conv3d(x, w, None)
    if extra16 is not None:
        e = extra16.float().permute(0, 4, 1, 2, 3).contiguous()
        y = y + F.conv3d(e, wx[:, :, None, None, None], None)
    if bias is not None:
        y = y + bias[None, :, None, None, None]
    y = y.permute(0, 2, 3, 4, 1).contiguous()
    if resid is not None:
        y = y + resid
    return y


def run_conv(nat, mode, B, D, H, W, cin, cout, cin_extra=0, terms=2, with_resid=False, impl=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    act = torch.randn(B, D, H, W, cin, device="cuda", generator=g).half()
    k = 1 if mode == 3 else 3
    w = torch.randn(cout, cin, k, k, k, device="cuda", generator=g) / (cin * k ** 3) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    if mode == 1:
        od, oh, ow = (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    elif mode == 2:
        od, oh, ow = 2 * D, 2 * H, 2 * W
    else:
        od, oh, ow = D, H, W
    extra = wx = None
    if cin_extra:
        extra = torch.randn(B, od, oh, ow, cin_extra, device="cuda", generator=g).half()
        wx = torch.randn(cout, cin_extra, device="cuda", generator=g) / cin_extra ** 0.5
    resid = torch.randn(B, od, oh, ow, cout, device="cuda", generator=g) if with_resid else None
    out32 = torch.full((B, od, oh, ow, cout), float("nan"), device="cuda")
    out16 = torch.zeros(B, od, oh, ow, cout, device="cuda", dtype=torch.half)
    rc = nat.lib().cm_op_conv3d(mode, nat.ptr(act), B, D, H, W, cin, nat.ptr(extra), cin_extra,
                                nat.ptr(w), nat.ptr(wx), nat.ptr(bias), cout, terms,
                                nat.ptr(resid), nat.ptr(out32), nat.ptr(out16), impl,
                                nat.current_stream())
    nat.check(rc)
    torch.cuda.synchronize()
    flag = nat.lib().cm_device_error()
    ref = torch_conv(mode, act, extra, w, wx, bias, resid)
    return out32, out16, ref, flag


CONV_CASES = [
    # (mode, B, D, H, W, cin, cout, cin_extra, resid)    -- names follow the ATC UNet layers
    (0, 2, 12, 36, 8, 32, 32, 0, True),     # encoder_blocks.0.conv_2 (identity residual), BK=32
    (0, 2, 12, 36, 8, 64, 32, 0, False),    # decoder_blocks.7.conv_1, BK=64 BN=32
    (0, 2, 6, 18, 4, 64, 64, 32, False),    # encoder_blocks.2.conv_2 + match_input slab, BK=32
    (0, 3, 3, 9, 2, 128, 128, 64, False),   # encoder_blocks.4.conv_2 + slab, partial last tile
    (0, 2, 3, 9, 2, 256, 128, 0, False),    # decoder_blocks.0.conv_1, BN=128
    (0, 1, 6, 18, 4, 192, 64, 0, False),    # decoder_blocks.3.conv_1
    (0, 1, 12, 36, 8, 96, 32, 0, False),    # decoder_blocks.6.conv_1, BK=32
    (1, 2, 12, 36, 8, 32, 32, 0, False),    # encoder_blocks.1.downsample (stride 2)
    (1, 3, 6, 18, 4, 64, 64, 0, False),     # encoder_blocks.3.downsample
    (2, 2, 3, 9, 2, 128, 128, 0, False),    # decoder_blocks.2.upsample (8 phase convs)
    (2, 1, 6, 18, 4, 64, 64, 0, False),     # decoder_blocks.5.upsample
    (3, 3, 3, 9, 2, 128, 384, 0, False),    # attention in_proj (N tiles = 3)
    (3, 3, 3, 9, 2, 128, 128, 0, True),     # attention out_proj + residual
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d_r%d" % c)
@pytest.mark.parametrize("terms", [2, 1])
def test_conv_umma_vs_torch(nat, case, terms):
    mode, B, D, H, W, cin, cout, cx, resid = case
    out32, out16, ref, flag = run_conv(nat, mode, B, D, H, W, cin, cout, cx, terms, resid, impl=0)
    assert flag == 0, f"device protocol error flag {flag}"
    assert not torch.isnan(out32).any(), "unwritten outputs: " + describe_mismatch(out32.nan_to_num(1e9), ref)
    e = rel_l2(out32, ref)
    bound = 2e-5 if terms == 2 else 1e-3
    assert e <= bound, f"rel-L2 {e:.3e} > {bound}: " + describe_mismatch(out32, ref)
    assert rel_l2(out16.float(), ref) <= 1e-3


@pytest.mark.parametrize("case", CONV_CASES[:4] + CONV_CASES[7:10],
                         ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d_r%d" % c)
def test_conv_umma_vs_scalar_restatement(nat, case):
    mode, B, D, H, W, cin, cout, cx, resid = case
    a32, _, _, flag = run_conv(nat, mode, B, D, H, W, cin, cout, cx, 2, resid, impl=0)
    b32, _, ref, _ = run_conv(nat, mode, B, D, H, W, cin, cout, cx, 2, resid, impl=1)
    assert flag == 0
    assert rel_l2(b32, ref) <= 2e-5, "scalar restatement itself is off: " + describe_mismatch(b32, ref)
    assert rel_l2(a32, b32) <= 1e-5, describe_mismatch(a32, b32)


@pytest.mark.parametrize("B,pixels,c0,c1,silu", [
    (2, 3456, 32, 0, 1), (2, 3456, 64, 32, 1), (3, 432, 128, 64, 1), (2, 54, 128, 0, 0),
    (2, 54, 128, 128, 1), (1, 5376, 64, 32, 1),
    # ragged slices of the streaming ring: pixel counts that do not divide by the slice / stage sizes
    (5, 777, 32, 64, 1), (2, 1000, 96, 0, 1), (3, 2305, 32, 0, 0), (64, 432, 64, 0, 1), (1, 20000, 32, 0, 1),
])
def test_gn_silu(nat, B, pixels, c0, c1, silu):
    g = torch.Generator(device="cuda").manual_seed(1)
    C_ = c0 + c1
    s0 = torch.randn(B, pixels, c0, device="cuda", generator=g) * 2 + 0.5
    s1 = torch.randn(B, pixels, c1, device="cuda", generator=g) if c1 else None
    gamma = torch.randn(C_, device="cuda", generator=g)
    beta = torch.randn(C_, device="cuda", generator=g)
    out = torch.zeros(B, pixels, C_, device="cuda", dtype=torch.half)
    raw = torch.zeros(B, pixels, C_, device="cuda", dtype=torch.half)
    nat.check(nat.lib().cm_op_gn_silu(nat.ptr(s0), c0, nat.ptr(s1), c1, nat.ptr(gamma), nat.ptr(beta),
                                      B, pixels, 1e-5, silu, nat.ptr(out), nat.ptr(raw),
                                      nat.current_stream()))
    torch.cuda.synchronize()
    x = torch.cat([s0, s1], dim=2) if c1 else s0
    ref = F.group_norm(x.permute(0, 2, 1).contiguous(), 8, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1)
    # fp16 output rounding: 2^-11 relative
    assert rel_l2(out.float(), ref) <= 4e-4
    assert (out.float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(raw, x.half())


@pytest.mark.parametrize("B,S,C_", [(3, 54, 128), (2, 108, 256), (2, 12, 128), (1, 84, 128), (3, 64, 32), (2, 8, 64), (1, 128, 96)])
def test_attn_core(nat, B, S, C_):
    g = torch.Generator(device="cuda").manual_seed(2)
    heads = 4
    qkv = torch.randn(B, S, 3 * C_, device="cuda", generator=g)
    ctx = torch.zeros(B, S, C_, device="cuda", dtype=torch.half)
    nat.check(nat.lib().cm_op_attn_core(nat.ptr(qkv), nat.ptr(ctx), B, S, C_, heads, nat.current_stream()))
    torch.cuda.synchronize()
    dh = C_ // heads
    q, k, v = [t.reshape(B, S, heads, dh).transpose(1, 2) for t in qkv.split(C_, dim=2)]
    p = torch.softmax((q @ k.transpose(-1, -2)) / dh ** 0.5, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(B, S, C_)
    # P, V and the output are fp16 MMA operands (2^-11 relative each); Q K^T is fp32-accurate (hi+lo)
    assert rel_l2(ctx.float(), ref) <= 7e-4


@pytest.mark.parametrize("B,S,C_", [(3, 54, 128), (2, 108, 256), (2, 84, 128), (2, 1, 128), (1, 7, 64), (2, 27, 32), (1, 128, 128),
                                    (2, 33, 256)])
def test_attn_core_backward_vs_torch_autograd(nat, B, S, C_):
    """Register-tiled shared-memory GEMM backward of softmax(Q K^T / sqrt(dh)) V (fp32 throughout) against autograd in
    float64: ragged S (not a multiple of 4, S = 1), dh = 8 ... 64, the largest tile (S = 128) and the ATC_medium shape."""
    g = torch.Generator(device="cuda").manual_seed(4)
    heads = 4
    dh = C_ // heads
    qkv = torch.randn(B, S, 3 * C_, device="cuda", generator=g)
    dctx = torch.randn(B, S, C_, device="cuda", generator=g)
    dqkv = torch.full((B, S, 3 * C_), float("nan"), device="cuda")
    nat.check(nat.lib().cm_op_attn_core_backward(nat.ptr(qkv), nat.ptr(dctx), nat.ptr(dqkv), B, S, C_, heads,
                                                 nat.current_stream()))
    torch.cuda.synchronize()
    x = qkv.double().requires_grad_(True)
    q, k, v = [t.reshape(B, S, heads, dh).transpose(1, 2) for t in x.split(C_, dim=2)]
    p = torch.softmax((q @ k.transpose(-1, -2)) / dh ** 0.5, dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(B, S, C_)
    ctx.backward(dctx.double())
    assert torch.isfinite(dqkv).all()
    for name, sl in (("dQ", slice(0, C_)), ("dK", slice(C_, 2 * C_)), ("dV", slice(2 * C_, 3 * C_))):
        e = rel_l2(dqkv[..., sl].double(), x.grad[..., sl])
        assert e <= 2e-6, f"{name} rel-L2 {e:.3e}"
    again = torch.empty_like(dqkv)
    nat.check(nat.lib().cm_op_attn_core_backward(nat.ptr(qkv), nat.ptr(dctx), nat.ptr(again), B, S, C_, heads,
                                                 nat.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(dqkv, again), "attention backward is not deterministic"


@pytest.mark.parametrize("B,S", [(2, 54), (3, 84), (1, 12), (2, 1), (2, 16), (1, 100), (2, 128)])
def test_attn_block_fused(nat, B, S):
    """Fused AttentionBlock (layers.py:5-18) vs torch GroupNorm + nn.MultiheadAttention in fp32."""
    C_, heads = 128, 4
    torch.manual_seed(11)
    x = torch.randn(B, S, C_, device="cuda") * 1.5 + 0.3
    gn = torch.nn.GroupNorm(8, C_).cuda()
    mha = torch.nn.MultiheadAttention(C_, heads, batch_first=True).cuda()
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.uniform_(-0.5, 0.5)
        mha.in_proj_bias.uniform_(-0.2, 0.2)
        mha.out_proj.bias.uniform_(-0.2, 0.2)
    out = torch.zeros(B, S, C_, device="cuda")
    out16 = torch.zeros(B, S, C_, device="cuda", dtype=torch.half)
    nat.check(nat.lib().cm_op_attn_block(nat.ptr(x), nat.ptr(gn.weight.data), nat.ptr(gn.bias.data),
                                         nat.ptr(mha.in_proj_weight.data), nat.ptr(mha.in_proj_bias.data),
                                         nat.ptr(mha.out_proj.weight.data), nat.ptr(mha.out_proj.bias.data),
                                         nat.ptr(out), nat.ptr(out16), B, S, C_, heads, 1e-5, nat.current_stream()))
    torch.cuda.synchronize()
    with torch.no_grad():
        h = gn(x.transpose(1, 2)).transpose(1, 2)            # GroupNorm over (C, tokens) per sample
        ref = x + mha(h, h, h, need_weights=False)[0]
    # fp16 MMA operands (h, P, V, ctx) with fp32 accumulation; the residual x is added in fp32
    assert rel_l2(out, ref) <= 5e-4
    assert rel_l2(out16.float(), ref) <= 1e-3
    assert bool(torch.isfinite(out).all())


@pytest.mark.parametrize("B,H,W,P,Fu,cin,cout", [(2, 12, 36, 5, 3, 3, 32), (1, 8, 12, 5, 3, 3, 64), (2, 4, 4, 2, 2, 4, 32)])
def test_first_conv(nat, B, H, W, P, Fu, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, cin, H, W, Fu, device="cuda", generator=g)
    past = torch.randn(B, cin, H, W, P, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    out = torch.zeros(B, P + Fu, H, W, cout, device="cuda")
    nat.check(nat.lib().cm_op_first_conv(nat.ptr(x), nat.ptr(past), nat.ptr(w), nat.ptr(b), nat.ptr(out),
                                         B, H, W, P, Fu, cin, cout, nat.current_stream()))
    torch.cuda.synchronize()
    ref = F.conv3d(torch.cat([past, x], dim=4), w, b, padding=1).permute(0, 4, 2, 3, 1)   # [B, L, H, W, C]
    assert rel_l2(out, ref) <= 1e-5


@pytest.mark.parametrize("B,H,W,P,Fu,cin,cout", [(2, 12, 36, 5, 3, 32, 3), (1, 8, 12, 8, 8, 64, 3), (2, 4, 4, 2, 2, 32, 4)])
def test_final_conv(nat, B, H, W, P, Fu, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(4)
    L = P + Fu
    act = torch.randn(B, L, H, W, cin, device="cuda", generator=g).half()   # internal [B, L, H, W, C]
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    eps = torch.zeros(B, cout, H, W, Fu, device="cuda")
    nat.check(nat.lib().cm_op_final_conv(nat.ptr(act), nat.ptr(w), nat.ptr(b), nat.ptr(eps), B, H, W, L, P,
                                         cin, cout, nat.current_stream()))
    torch.cuda.synchronize()
    ref = F.conv3d(act.float().permute(0, 4, 2, 3, 1), w, b, padding=1)[..., P:]
    assert rel_l2(eps, ref) <= 1e-5


# ---------------------------------------------------------------------------------------------
# backward: data gradient (conv_umma with re-packed weights) and weight gradient (wgrad_umma,
# MN-major tcgen05) of every conv shape, against torch autograd on the same fp16-rounded operands
# ---------------------------------------------------------------------------------------------
def _out_grid(mode, D, H, W):
    if mode == 1:
        return (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if mode == 2:
        return 2 * D, 2 * H, 2 * W
    return D, H, W


def torch_conv_grads(mode, act16, extra16, w, wx, dout16):
    x = act16.float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    w = w.clone().requires_grad_(True)
    if mode == 0:
        y = F.conv3d(x, w, None, stride=1, padding=1)
    elif mode == 1:
        y = F.conv3d(x, w, None, stride=2, padding=1)
    elif mode == 2:
        y = F.conv3d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, stride=1, padding=1)
    else:
        y = F.conv3d(x, w, None)
    wxg = None
    if extra16 is not None:
        e = extra16.float().permute(0, 4, 1, 2, 3).contiguous()
        wxg = wx.clone().requires_grad_(True)
        y = y + F.conv3d(e, wxg[:, :, None, None, None], None)
    y.backward(dout16.float().permute(0, 4, 1, 2, 3).contiguous())
    dx = x.grad.permute(0, 2, 3, 4, 1).contiguous()
    return dx, w.grad, (wxg.grad if wxg is not None else None)


BWD_CASES = [
    # (mode, B, D, H, W, cin, cout, cin_extra)
    (0, 2, 12, 36, 8, 32, 32, 0),
    (0, 2, 6, 18, 4, 64, 64, 32),
    (0, 3, 3, 9, 2, 128, 128, 64),
    (0, 2, 3, 9, 2, 256, 128, 0),
    (0, 1, 12, 36, 8, 96, 32, 0),
    (0, 1, 6, 18, 4, 192, 64, 0),
    (1, 2, 12, 36, 8, 32, 32, 0),
    (1, 3, 6, 18, 4, 64, 64, 0),
    (2, 2, 3, 9, 2, 128, 128, 0),
    (2, 1, 6, 18, 4, 64, 64, 0),
    (3, 3, 3, 9, 2, 128, 384, 0),
    (3, 3, 3, 9, 2, 128, 128, 0),
]


def _bwd_inputs(mode, B, D, H, W, cin, cout, cx, seed=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    act = torch.randn(B, D, H, W, cin, device="cuda", generator=g).half()
    k = 1 if mode == 3 else 3
    w = torch.randn(cout, cin, k, k, k, device="cuda", generator=g) / (cin * k ** 3) ** 0.5
    od, oh, ow = _out_grid(mode, D, H, W)
    extra = wx = None
    if cx:
        extra = torch.randn(B, od, oh, ow, cx, device="cuda", generator=g).half()
        wx = torch.randn(cout, cx, device="cuda", generator=g) / cx ** 0.5
    dout = torch.randn(B, od, oh, ow, cout, device="cuda", generator=g).half()
    return act, w, extra, wx, dout


@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d" % c)
def test_conv_dgrad_vs_torch_autograd(nat, case):
    mode, B, D, H, W, cin, cout, cx = case
    act, w, extra, wx, dout = _bwd_inputs(*case)
    dx = torch.full((B, D, H, W, cin), float("nan"), device="cuda")
    nat.check(nat.lib().cm_op_conv3d_dgrad(mode, nat.ptr(dout), B, D, H, W, cin, nat.ptr(w), cout, 2,
                                           nat.ptr(dx), nat.current_stream()))
    torch.cuda.synchronize()
    assert nat.lib().cm_device_error() == 0
    ref, _, _ = torch_conv_grads(mode, act, None, w, None, dout)
    e = rel_l2(dx, ref)
    assert e <= 3e-5, f"dgrad rel-L2 {e:.3e}; " + describe_mismatch(dx, ref)


@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d" % c)
@pytest.mark.parametrize("dup", [2, 1], ids=["hilo", "single"])
def test_conv_dgrad_from_fp32_gradient_hi_lo_pair(nat, case, dup):
    """The training backward's dgrad operand: an fp32 gradient cast to fp16 (dup = 1: rounding error ~2.8e-4
    per element) or to a K-concatenated hi|lo fp16 pair (dup = 2: exact to 2^-22), against torch autograd on
    the UNROUNDED fp32 gradient.  Bounds: 2e-5 (hi|lo: fp32 accumulation order only) / 1e-3 (single)."""
    mode, B, D, H, W, cin, cout, cx = case
    act, w, _, _, _ = _bwd_inputs(*case)
    od, oh, ow = _out_grid(mode, D, H, W)
    g = torch.Generator(device="cuda").manual_seed(17)
    dout32 = torch.randn(B, od, oh, ow, cout, device="cuda", generator=g) * 3.0
    dx = torch.full((B, D, H, W, cin), float("nan"), device="cuda")
    nat.check(nat.lib().cm_op_conv3d_dgrad_f32(mode, nat.ptr(dout32), B, D, H, W, cin, nat.ptr(w), cout, 2, dup,
                                               nat.ptr(dx), nat.current_stream()))
    torch.cuda.synchronize()
    assert nat.lib().cm_device_error() == 0
    ref, _, _ = torch_conv_grads(mode, act, None, w, None, dout32)
    e = rel_l2(dx, ref)
    print(f"dgrad from fp32 dOut, dup={dup}: rel-L2 {e:.3e}")
    assert e <= (2e-5 if dup == 2 else 1e-3), f"dgrad rel-L2 {e:.3e}; " + describe_mismatch(dx, ref)


@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d" % c)
@pytest.mark.parametrize("impl", [1, 0], ids=["scalar", "umma"])
def test_conv_wgrad_vs_torch_autograd(nat, case, impl):
    mode, B, D, H, W, cin, cout, cx = case
    act, w, extra, wx, dout = _bwd_inputs(*case)
    dw = torch.full_like(w, float("nan"))
    dwx = torch.full_like(wx, float("nan")) if cx else None
    nat.check(nat.lib().cm_op_conv3d_wgrad(mode, nat.ptr(act), B, D, H, W, cin, nat.ptr(extra), cx,
                                           nat.ptr(dout), cout, nat.ptr(dw), nat.ptr(dwx), impl,
                                           nat.current_stream()))
    torch.cuda.synchronize()
    assert nat.lib().cm_device_error() == 0
    _, rw, rwx = torch_conv_grads(mode, act, extra, w, wx, dout)
    e = rel_l2(dw, rw)
    assert e <= 3e-5, f"wgrad rel-L2 {e:.3e}; " + describe_mismatch(dw.reshape(cout, -1), rw.reshape(cout, -1))
    if cx:
        assert rel_l2(dwx, rwx) <= 3e-5


# plane / halo weight gradient (wgrad_plane.cuh): one haloed box per td plane + one dOut box per unit serve all 27 taps
# (th = row-offset views, tw = the M dimension's chunks one row apart), nine accumulators resident in TMEM.  impl=2 forces it.
WGRAD_PLANE_CASES = [
    (0, 2, 8, 12, 36, 32, 32, 0),      # encoder_blocks.0.conv_1 / conv_2 (ATC): ragged last row block (12 = 8 + 4)
    (0, 1, 8, 28, 24, 32, 32, 0),      # HERMES-CR-120 level 0: 4 row blocks, last one half empty
    (0, 3, 8, 8, 12, 32, 32, 0),       # ETH-UCY level 0: whole planes
    (0, 2, 8, 12, 36, 96, 32, 0),      # decoder_blocks.6.conv_1: three 32-channel chunks
    (0, 2, 8, 12, 36, 32, 32, 96),     # decoder_blocks.6.conv_2 + match_input rows (1x1x1 source through wgrad_umma)
    (0, 40, 8, 12, 36, 64, 32, 0),     # more units than SMs: several units accumulate in one CTA's TMEM
    (0, 1, 2, 5, 9, 32, 32, 0),        # odd width (W + 2 = 11): 16-row blocks, H < HB
]


@pytest.mark.parametrize("case", WGRAD_PLANE_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d" % c)
def test_conv_wgrad_plane_vs_torch_autograd(nat, case):
    mode, B, D, H, W, cin, cout, cx = case
    act, w, extra, wx, dout = _bwd_inputs(*case)
    dw = torch.full_like(w, float("nan"))
    dwx = torch.full_like(wx, float("nan")) if cx else None
    nat.check(nat.lib().cm_op_conv3d_wgrad(mode, nat.ptr(act), B, D, H, W, cin, nat.ptr(extra), cx,
                                           nat.ptr(dout), cout, nat.ptr(dw), nat.ptr(dwx), 2,
                                           nat.current_stream()))
    torch.cuda.synchronize()
    assert nat.lib().cm_device_error() == 0
    _, rw, rwx = torch_conv_grads(mode, act, extra, w, wx, dout)
    e = rel_l2(dw, rw)
    assert e <= 3e-5, f"plane wgrad rel-L2 {e:.3e}; " + describe_mismatch(dw.reshape(cout, -1), rw.reshape(cout, -1))
    if cx:
        assert rel_l2(dwx, rwx) <= 3e-5


# ---------------------------------------------------------------------------------------------
# plane-tile conv (conv_plane.cuh): W-shifted descriptor views of one haloed TMA box, stacked hi|lo
# MMA, persistent CTAs with double-buffered TMEM.  impl=2 forces it (error if not covered).
# ---------------------------------------------------------------------------------------------
PLANE_CASES = [
    (0, 2, 8, 12, 36, 32, 32, 0, True),     # encoder_blocks.0.conv_2 (HB=6 row blocks)
    (0, 2, 8, 12, 36, 64, 32, 0, False),    # decoder_blocks.7.conv_1 (BK=64)
    (0, 1, 8, 12, 36, 96, 32, 0, False),    # decoder_blocks.6.conv_1 (3 channel chunks)
    (0, 2, 8, 12, 36, 32, 32, 96, False),   # decoder_blocks.6.conv_2 + match_input slab
    (0, 3, 4, 6, 18, 64, 64, 32, False),    # encoder_blocks.2.conv_2 + slab, odd batch
    (0, 1, 4, 6, 18, 192, 64, 0, False),    # decoder_blocks.3.conv_1 (R=2 planes per unit)
    (0, 1, 8, 28, 24, 32, 32, 0, True),     # HERMES-CR-120 level 0
    (0, 5, 8, 8, 12, 32, 32, 0, True),      # ETH-UCY level 0
    (0, 150, 8, 12, 36, 32, 32, 0, True),   # more units than SMs: several units per persistent CTA
    (0, 2, 8, 12, 36, 96, 32, 96, True),    # th3 stages: 3 chunks per td + 3 match_input slabs + residual
    (0, 2, 4, 6, 18, 128, 64, 64, False),   # th3 stages at level 1: 4 chunks per td + 2 match_input slabs
    (0, 3, 3, 12, 36, 64, 32, 0, True),     # odd depth and batch, 2 chunks per td, residual
]


@pytest.mark.parametrize("case", PLANE_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d_r%d" % c)
@pytest.mark.parametrize("terms", [2, 1])
def test_conv_plane_vs_torch(nat, case, terms):
    mode, B, D, H, W, cin, cout, cx, resid = case
    out32, out16, ref, flag = run_conv(nat, mode, B, D, H, W, cin, cout, cx, terms, resid, impl=2)
    assert flag == 0, f"device protocol error flag {flag}"
    e = rel_l2(out32, ref)
    assert e <= (2e-5 if terms == 2 else 1e-3), f"rel-L2 {e:.3e}; " + describe_mismatch(out32, ref)
    assert rel_l2(out16.float(), ref) <= 1e-3


# ---------------------------------------------------------------------------------------------
# weights-resident 32 -> 32 conv (conv_res32.cuh): tw taps and hi|lo terms stacked along N = 192, one-tile
# units, weights loaded once per CTA, drain / store warp pipeline.  impl=3 forces it (error if not covered).
# ---------------------------------------------------------------------------------------------
RES32_CASES = [
    (0, 2, 8, 12, 36, 32, 32, 0, True),     # encoder_blocks.0.conv_2 (identity residual), HB = 3
    (0, 2, 8, 12, 36, 32, 32, 0, False),    # encoder_blocks.0.conv_1 / UNet.first
    (0, 2, 8, 12, 36, 32, 32, 96, False),   # decoder_blocks.6.conv_2 + match_input (3 chunks)
    (0, 3, 8, 12, 36, 32, 32, 64, False),   # decoder_blocks.7.conv_2 + match_input (2 chunks), odd batch
    (0, 1, 8, 28, 24, 32, 32, 0, True),     # HERMES-CR-120 level 0 (HB = 4)
    (0, 5, 8, 8, 12, 32, 32, 32, True),     # ETH-UCY level 0 (whole plane per unit) + slab + residual
    (0, 150, 8, 12, 36, 32, 32, 0, True),   # many units per persistent CTA (both TMEM / transpose buffers reused)
    (0, 1, 3, 12, 36, 32, 32, 0, False),    # fewer units than SMs, odd depth
]


@pytest.mark.parametrize("case", RES32_CASES, ids=lambda c: "m%d_B%d_%dx%dx%d_ci%d_co%d_x%d_r%d" % c)
def test_conv_res32_vs_torch(nat, case):
    mode, B, D, H, W, cin, cout, cx, resid = case
    out32, out16, ref, flag = run_conv(nat, mode, B, D, H, W, cin, cout, cx, 2, resid, impl=3)
    assert flag == 0, f"device protocol error flag {flag}"
    e = rel_l2(out32, ref)
    assert e <= 2e-5, f"rel-L2 {e:.3e}; " + describe_mismatch(out32, ref)
    assert rel_l2(out16.float(), ref) <= 1e-3


def test_conv_res32_matches_plane_kernel_and_is_deterministic(nat):
    """Same operands through conv_plane_kernel and conv_res32_kernel agree to fp32 summation order; two runs of
    the resident kernel are bit-identical."""
    a, _, _, _ = run_conv(nat, 0, 4, 8, 12, 36, 32, 32, 96, 2, True, impl=3)
    b, _, _, _ = run_conv(nat, 0, 4, 8, 12, 36, 32, 32, 96, 2, True, impl=3)
    c, _, _, _ = run_conv(nat, 0, 4, 8, 12, 36, 32, 32, 96, 2, True, impl=2)
    assert torch.equal(a, b)
    assert rel_l2(a, c) <= 2e-6
