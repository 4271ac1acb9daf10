"""TEST INFRASTRUCTURE — CPU restatement of the reference UNet forward (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; the product package never does (it has no CPU path).

A functional, state_dict-driven restatement of
  models/backbones/unet.py:124-167      UNet.forward
  models/backbones/layers.py:55-78      ResnetBlock.forward
  models/backbones/layers.py:12-18      AttentionBlock.forward
  models/backbones/layers.py:81-96      DownSample / UpSample
  models/backbones/embeddings.py:33-34  SinusoidalPositionEmbeddings.forward
with the arithmetic done by the same third-party library the reference calls
(torch.nn.functional on CPU: torch==2.5.1 pinned by the reference's requirements.txt:14,
2.11.0 in this image).  Pinned against the live reference in tests/test_oracle_pins.py and
against the committed vectors in tests/golden/ (generated from the reference by
oracle/make_golden.py).  The reference ships NO golden vectors of its own.

`operand_dtype` emulates the CUDA path's operand rounding (fp16 activations into every
conv / projection, optional hi+lo split weights) to predict its error budget on CPU.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def sinusoid_table(total_time_steps: int, dim: int, dtype=torch.float32) -> Tensor:
    """embeddings.py:10-20 — frozen [T, dim] table (sin | cos)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    ang = torch.arange(total_time_steps, dtype=torch.float32)[:, None] * freq[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1).to(dtype)


class _Emu:
    """Operand-rounding emulation of the tcgen05 path (None = exact fp32/fp64 oracle)."""

    def __init__(self, operand_dtype: Optional[torch.dtype], weight_terms: int):
        self.dt = operand_dtype
        self.terms = weight_terms

    def act(self, x: Tensor) -> Tensor:
        return x if self.dt is None else x.to(self.dt).to(x.dtype)

    def weight(self, w: Tensor) -> Tensor:
        if self.dt is None:
            return w
        hi = w.to(self.dt).to(w.dtype)
        if self.terms == 1:
            return hi
        return hi + (w - hi).to(self.dt).to(w.dtype)


def _gn(x: Tensor, sd, prefix: str) -> Tensor:
    return F.group_norm(x, 8, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _attention(x: Tensor, sd, prefix: str, emu: _Emu) -> Tensor:
    """layers.py:12-18 with nn.MultiheadAttention(embed_dim=C, num_heads=4, batch_first=True)."""
    B, C, H, W, L = x.shape
    heads, dh = 4, C // 4
    h = _gn(x, sd, prefix + ".group_norm")
    h = h.reshape(B, C, H * W * L).swapaxes(1, 2)                       # [B, S, C]
    qkv = F.linear(emu.act(h), emu.weight(sd[prefix + ".mhsa.in_proj_weight"]),
                   sd[prefix + ".mhsa.in_proj_bias"])
    q, k, v = qkv.split(C, dim=2)
    q = q.reshape(B, -1, heads, dh).transpose(1, 2)
    k = k.reshape(B, -1, heads, dh).transpose(1, 2)
    v = v.reshape(B, -1, heads, dh).transpose(1, 2)
    p = torch.softmax((q / math.sqrt(dh)) @ k.transpose(-1, -2), dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(B, -1, C)
    out = F.linear(emu.act(ctx), emu.weight(sd[prefix + ".mhsa.out_proj.weight"]),
                   sd[prefix + ".mhsa.out_proj.bias"])
    out = out.swapaxes(2, 1).reshape(B, C, H, W, L)
    return x + out


def _resblock(x: Tensor, temb: Tensor, sd, prefix: str, emu: _Emu,
              drop_masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """layers.py:55-78."""
    h = F.silu(_gn(x, sd, prefix + ".normalize_1"))
    h = F.conv3d(emu.act(h), emu.weight(sd[prefix + ".conv_1.weight"]), sd[prefix + ".conv_1.bias"],
                 padding=1)
    h = h + F.linear(F.silu(temb), sd[prefix + ".dense_1.weight"],
                     sd[prefix + ".dense_1.bias"])[:, :, None, None, None]
    h = F.silu(_gn(h, sd, prefix + ".normalize_2"))
    if drop_masks is not None and prefix in drop_masks:
        h = h * drop_masks[prefix][:, :, None, None, None]            # Dropout3d (layers.py:71)
    h = F.conv3d(emu.act(h), emu.weight(sd[prefix + ".conv_2.weight"]), sd[prefix + ".conv_2.bias"],
                 padding=1)
    if prefix + ".match_input.weight" in sd:
        sc = F.conv3d(emu.act(x), emu.weight(sd[prefix + ".match_input.weight"]),
                      sd[prefix + ".match_input.bias"])
    else:
        sc = x
    h = h + sc
    if prefix + ".attention.group_norm.weight" in sd:
        h = _attention(h, sd, prefix + ".attention", emu)
    return h


def unet_forward(sd: Dict[str, Tensor], future: Tensor, t: Tensor, past: Tensor, *,
                 num_res_blocks: int, num_levels: int,
                 operand_dtype: Optional[torch.dtype] = None, weight_terms: int = 2,
                 drop_masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """eps = UNet(...).forward(future, t, past) (unet.py:124-167), from a reference state_dict.

    The block structure is recovered from `num_res_blocks` / `num_levels` exactly as
    UNet.__init__ lays it out (unet.py:40-115); attention / match_input presence is read off
    the state_dict keys.
    """
    emu = _Emu(operand_dtype, weight_terms)
    P = past.shape[4]
    temb = F.embedding(t, sd["time_embeddings.time_blocks.0.weight"])
    temb = F.linear(temb, sd["time_embeddings.time_blocks.1.weight"], sd["time_embeddings.time_blocks.1.bias"])
    temb = F.linear(F.silu(temb), sd["time_embeddings.time_blocks.3.weight"],
                    sd["time_embeddings.time_blocks.3.bias"])
    x = torch.cat([past, future], dim=4)                                   # unet.py:138
    h = F.conv3d(x, sd["first.weight"], sd["first.bias"], padding=1)       # unet.py:144
    outs = [h]
    idx = 0
    for level in range(num_levels):                                       # unet.py:148-150
        for _ in range(num_res_blocks):
            h = _resblock(h, temb, sd, f"encoder_blocks.{idx}", emu, drop_masks)
            outs.append(h)
            idx += 1
        if level != num_levels - 1:
            p = f"encoder_blocks.{idx}.downsample"
            h = F.conv3d(emu.act(h), emu.weight(sd[p + ".weight"]), sd[p + ".bias"], stride=2, padding=1)
            outs.append(h)
            idx += 1
    for i in range(2):                                                     # unet.py:153-154
        h = _resblock(h, temb, sd, f"bottleneck_blocks.{i}", emu, drop_masks)
    idx = 0
    for level in reversed(range(num_levels)):                              # unet.py:157-161
        for _ in range(num_res_blocks + 1):
            h = torch.cat([h, outs.pop()], dim=1)
            h = _resblock(h, temb, sd, f"decoder_blocks.{idx}", emu, drop_masks)
            idx += 1
        if level != 0:
            p = f"decoder_blocks.{idx}.upsample.1"
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv3d(emu.act(h), emu.weight(sd[p + ".weight"]), sd[p + ".bias"], padding=1)
            idx += 1
    h = F.silu(_gn(h, sd, "final.0"))                                      # unet.py:118-122,163
    h = F.conv3d(emu.act(h), sd["final.2.weight"], sd["final.2.bias"], padding=1)
    return h[..., P:]                                                      # unet.py:165-167


def rel_l2(a: Tensor, b: Tensor) -> float:
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
