"""Forward diffusion process — drop-in for reference models/diffusion/forward.py:1-37.

Same class, buffers (``beta, alpha, alpha_bar, sqrt_alpha_bar, one_by_sqrt_alpha,
sqrt_one_minus_alpha_bar``), ``timesteps`` attribute and ``forward(x0, t) -> (x_t, eps)``
contract.  The schedule is host-side set-up (six [T] fp32 vectors); its values feed the
per-step coefficient table consumed by the fused reverse-step epilogue (cm_ddpm_sample).
"""
import torch
import torch.nn as nn


def get_from_idx(element: torch.Tensor, idx: torch.Tensor):
    """Gather per-sample schedule entries and make them broadcastable over [B,C,H,W,L]."""
    return element.gather(-1, idx).reshape(-1, 1, 1, 1, 1)


class ForwardSampler(nn.Module):
    def __init__(self, timesteps=1000, scale=1, beta_start=1e-4, beta_end=2e-2):
        super().__init__()
        self.timesteps = timesteps
        beta = torch.linspace(scale * beta_start, scale * beta_end, timesteps, dtype=torch.float32)
        alpha = 1 - beta
        alpha_bar = torch.cumprod(alpha, dim=0)
        for name, value in (("beta", beta), ("alpha", alpha), ("alpha_bar", alpha_bar),
                            ("sqrt_alpha_bar", torch.sqrt(alpha_bar)),
                            ("one_by_sqrt_alpha", 1. / torch.sqrt(alpha)),
                            ("sqrt_one_minus_alpha_bar", torch.sqrt(1 - alpha_bar))):
            self.register_buffer(name, value)

    def forward(self, x0: torch.Tensor, timesteps: torch.Tensor):
        """q(x_t | x_0): returns (x_t, eps) with eps ~ N(0, I) from torch's generator."""
        epsilon = torch.randn_like(x0)
        mean = get_from_idx(self.sqrt_alpha_bar, timesteps) * x0
        std_dev = get_from_idx(self.sqrt_one_minus_alpha_bar, timesteps)
        return mean + std_dev * epsilon, epsilon
