"""TEST INFRASTRUCTURE — CPU restatement (torch functional, fp32 or fp64) of the reference's second noise-prediction
backbone, `DiT4D_V4`.  Only tests/, smoke() and bench.py's cpu_baseline leg may import this;
the product path (crowdmod-ddpm-4d_b200/models/backbones/DiT4D_V4.py -> cm_dit_forward / cm_dit_sample) never does.

Follows /root/reference/models/backbones/DiT4D_V4.py:
  * patch_embed        <- PatchEmbed4D.forward          (:47-63)   Conv3d with kernel = stride = (t_patch, p, p)
  * add_pos            <- _add_positional_embeddings    (:329-346)
  * block              <- DiTBlockCA.forward            (:143-210) spatial self-attention per temporal slot, temporal
                          cross-attention (future slots query all slots), MLP, all AdaLN-Zero modulated and gated
  * final / unpatch    <- FinalLayer.forward (:232-234), PatchUnEmbed4D.forward (:80-102)
  * dit_forward        <- DiT4D_V4.forward              (:348-375)
and /root/reference/models/backbones/embeddings.py:33-34 for the diffusion-time embedding.
`sd` is the module's state_dict (reference key names).  Pinned against the live reference in tests/test_oracle_pins.py
and against golden vectors generated from it (oracle/make_golden.py -> tests/golden/dit_*.npz).
"""
import torch
import torch.nn.functional as F


def _mha(q_in, kv_in, w_in, b_in, w_out, b_out, heads):
    """nn.MultiheadAttention(batch_first=True, eval) with separate query / key=value inputs."""
    D = q_in.shape[-1]
    dh = D // heads
    q = F.linear(q_in, w_in[:D], b_in[:D])
    k = F.linear(kv_in, w_in[D:2 * D], b_in[D:2 * D])
    v = F.linear(kv_in, w_in[2 * D:], b_in[2 * D:])
    Bq, Sq, _ = q.shape
    Sk = k.shape[1]
    q = q.reshape(Bq, Sq, heads, dh).transpose(1, 2)
    k = k.reshape(Bq, Sk, heads, dh).transpose(1, 2)
    v = v.reshape(Bq, Sk, heads, dh).transpose(1, 2)
    p = torch.softmax((q @ k.transpose(-1, -2)) / dh ** 0.5, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(Bq, Sq, D)
    return F.linear(o, w_out, b_out)


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _mod(x, shift, scale):
    return x * (1.0 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def time_condition(sd, t):
    """c = SiLU(time_proj(time_blocks(t)))  (DiT4D_V4.py:363; embeddings.py:22-34)."""
    p = "dif_time_embeddings.time_blocks."
    e = sd[p + "0.weight"][t]
    h = F.silu(F.linear(e, sd[p + "1.weight"], sd[p + "1.bias"]))
    h = F.linear(h, sd[p + "3.weight"], sd[p + "3.bias"])
    return F.silu(F.linear(h, sd["time_proj.0.weight"], sd["time_proj.0.bias"]))


def block(sd, pre, x, c, heads, Ns, Tp, qs):
    B, _, D = x.shape
    m = F.linear(F.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
    sh1, sc1, g1, sh2, sc2, g2, sh3, sc3, g3 = m.chunk(9, dim=-1)
    # 1. spatial self-attention: every temporal slot is its own sequence of Ns tokens
    xs = x.reshape(B * Tp, Ns, D)
    rep = lambda v, k: v.repeat_interleave(k, dim=0)
    xm = _mod(_ln(xs), rep(sh1, Tp), rep(sc1, Tp))
    a = _mha(xm, xm, sd[pre + "spatial_attn.in_proj_weight"], sd[pre + "spatial_attn.in_proj_bias"],
             sd[pre + "spatial_attn.out_proj.weight"], sd[pre + "spatial_attn.out_proj.bias"], heads)
    x = (xs + rep(g1, Tp).unsqueeze(1) * a).reshape(B, Tp * Ns, D)
    # 2. temporal cross-attention: every spatial patch is its own sequence of Tp slots; future slots query all
    xt = x.reshape(B, Tp, Ns, D).permute(0, 2, 1, 3).reshape(B * Ns, Tp, D)
    kv = _mod(_ln(xt), rep(sh2, Ns), rep(sc2, Ns))
    a = _mha(kv[:, qs:], kv, sd[pre + "temporal_attn.in_proj_weight"], sd[pre + "temporal_attn.in_proj_bias"],
             sd[pre + "temporal_attn.out_proj.weight"], sd[pre + "temporal_attn.out_proj.bias"], heads)
    xt = torch.cat([xt[:, :qs], xt[:, qs:] + rep(g2, Ns).unsqueeze(1) * a], dim=1)
    x = xt.reshape(B, Ns, Tp, D).permute(0, 2, 1, 3).reshape(B, Tp * Ns, D)
    # 3. MLP (exact erf GELU, nn.GELU default)
    xm = _mod(_ln(x), sh3, sc3)
    h = F.linear(F.gelu(F.linear(xm, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"])),
                 sd[pre + "mlp.3.weight"], sd[pre + "mlp.3.bias"])
    return x + g3.unsqueeze(1) * h


def dit_forward(sd, future, t, past, *, patch, t_patch, heads, depth):
    """future [B, C, H, W, F], t [B] int64, past [B, C, H, W, P] -> predicted noise [B, C, H, W, F] (eval mode)."""
    x = torch.cat([past, future], dim=4)
    B, C, H, W, T = x.shape
    P = past.shape[4]
    D = sd["patch_embed.proj.weight"].shape[0]
    c = time_condition(sd, t)
    tok = F.conv3d(x.permute(0, 1, 4, 2, 3), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"],
                   stride=(t_patch, patch, patch))                  # (B, D, Tp, hp, wp)
    Tp, hp, wp = tok.shape[2:]
    Ns = hp * wp
    tok = tok.permute(0, 2, 3, 4, 1).reshape(B, Tp, Ns, D)
    tok = tok + sd["spatial_pos_embed"].unsqueeze(1) + sd["temporal_pos_embed"][:, :Tp].unsqueeze(2)
    tok = tok.reshape(B, Tp * Ns, D)
    qs = P // t_patch
    for i in range(depth):
        tok = block(sd, f"blocks.{i}.", tok, c, heads, Ns, Tp, qs)
    m = F.linear(F.silu(c), sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = m.chunk(2, dim=-1)
    out = F.linear(_mod(_ln(tok), shift, scale), sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    Co = out.shape[-1] // (t_patch * patch * patch)
    out = out.reshape(B, Tp, hp, wp, t_patch, Co, patch, patch).permute(0, 5, 1, 4, 2, 6, 3, 7)
    out = out.reshape(B, Co, Tp * t_patch, hp * patch, wp * patch).permute(0, 1, 3, 4, 2)
    return out[:, :, :, :, P:]


def randomize_zero_init(sd, seed):
    """The reference zero-initialises the AdaLN-Zero and final projections (DiT4D_V4.py:139-140, :226-229): a freshly
    constructed model outputs exactly 0.  Parity tests fill those tensors with small seeded noise (what training
    would do) so that every path of the block is exercised."""
    g = torch.Generator().manual_seed(seed)
    out = dict(sd)
    for k, v in sd.items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            out[k] = (0.05 * torch.randn(v.shape, generator=g)).to(v.dtype)
    return out
