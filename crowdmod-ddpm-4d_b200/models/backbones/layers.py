"""Backbone building blocks (mirror of reference models/backbones/layers.py:5-96).

Same constructor signatures, attribute names and therefore the same state_dict keys and the
same default initialisation (parameters are created by the same torch.nn constructors in the
same order, so ``torch.manual_seed(s); UNet(...)`` reproduces the reference's weights bit for
bit).  The blocks are parameter containers: their arithmetic is fused across block boundaries
in the native plan (GroupNorm+SiLU -> fp16 operand, conv epilogues carrying bias / time
embedding / residual, match_input as a K-slab of conv_2, upsample folded into the conv), so
they are not individually callable.
"""
import torch.nn as nn


def _fused(name):
    def forward(self, *args, **kwargs):
        raise RuntimeError(f"{name} is fused into the native UNet plan; call UNet.forward")
    return forward


class AttentionBlock(nn.Module):
    def __init__(self, channels=64):
        super().__init__()
        self.channels = channels
        self.group_norm = nn.GroupNorm(num_groups=8, num_channels=channels)
        self.mhsa = nn.MultiheadAttention(embed_dim=channels, num_heads=4, batch_first=True)

    forward = _fused("AttentionBlock")


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels, dropout_rate=0.1, time_emb_dims=512,
                 apply_attention=False, condition="Past"):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.condition = condition
        self.activation = nn.SiLU()
        self.normalize_1 = nn.GroupNorm(num_groups=8, num_channels=in_channels)
        self.conv_1 = nn.Conv3d(in_channels, out_channels, kernel_size=3, stride=1, padding="same")
        self.dense_1 = nn.Linear(time_emb_dims, out_channels)
        self.normalize_2 = nn.GroupNorm(num_groups=8, num_channels=out_channels)
        self.dropout = nn.Dropout3d(p=dropout_rate)
        self.conv_2 = nn.Conv3d(out_channels, out_channels, kernel_size=3, stride=1, padding="same")
        if in_channels != out_channels:
            self.match_input = nn.Conv3d(in_channels, out_channels, kernel_size=1, stride=1)
        else:
            self.match_input = nn.Identity()
        self.attention = AttentionBlock(channels=out_channels) if apply_attention else nn.Identity()

    forward = _fused("ResnetBlock")


class DownSample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.downsample = nn.Conv3d(channels, channels, kernel_size=3, stride=2, padding=1)

    forward = _fused("DownSample")


class UpSample(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.upsample = nn.Sequential(
            nn.Upsample(scale_factor=2, mode="nearest"),
            nn.Conv3d(in_channels, in_channels, kernel_size=3, stride=1, padding=1))

    forward = _fused("UpSample")
