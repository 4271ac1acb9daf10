"""Time-step embedding parameters (mirror of reference models/backbones/embeddings.py:6-34).

Holds exactly the reference's state_dict entries
``time_blocks.{0.weight[1000,d] (frozen sinusoid table), 1.weight, 1.bias, 3.weight, 3.bias}``.
The arithmetic runs in the native library (csrc/kernels.cu temb_kernel), fused with every
ResnetBlock's dense_1 projection; this module is a parameter container.
"""
import math

import torch
import torch.nn as nn


def _sinusoid_table(total_time_steps: int, dims: int) -> torch.Tensor:
    half = dims // 2
    scale = math.log(10000) / (half - 1)
    freqs = torch.exp(-scale * torch.arange(half, dtype=torch.float32))
    angles = torch.arange(total_time_steps, dtype=torch.float32).unsqueeze(-1) * freqs.unsqueeze(0)
    return torch.cat((angles.sin(), angles.cos()), dim=-1)


class SinusoidalPositionEmbeddings(nn.Module):
    def __init__(self, total_time_steps=1000, time_emb_dims=128, time_emb_dims_exp=512):
        super().__init__()
        self.total_time_steps = total_time_steps
        self.time_blocks = nn.Sequential(
            nn.Embedding.from_pretrained(_sinusoid_table(total_time_steps, time_emb_dims)),
            nn.Linear(time_emb_dims, time_emb_dims_exp),
            nn.SiLU(),
            nn.Linear(time_emb_dims_exp, time_emb_dims_exp),
        )

    def forward(self, time):
        raise RuntimeError(
            "SinusoidalPositionEmbeddings is evaluated inside the fused native UNet plan "
            "(cm_unet_forward); call UNet.forward instead")
