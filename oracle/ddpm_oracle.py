"""TEST INFRASTRUCTURE — CPU restatement of the reference diffusion math (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; the product package never does.

Restates, with injectable noise instead of the global torch generator,
  models/diffusion/forward.py:10-27   ForwardSampler.__init__  (schedule buffers)
  models/diffusion/forward.py:29-37   ForwardSampler.forward   (q(x_t | x_0))
  models/diffusion/ddpm.py:25-38      DDPM.step
  models/diffusion/ddpm.py:206-236    DDPM_model._generate_ddpm (incl. Sparsity guidance :223-226)
  models/diffusion/ddpm.py:238-282    DDPM_model._generate_ddim
  models/diffusion/ddpm.py:111-121    DDPM_model._train_step    (loss)
Noise order of the reference: x_T first, then one randn_like per step for
t = T-1 .. 1 (none at t = 0); DDIM draws one per step including the last.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def schedule(timesteps: int = 1000, scale: float = 1.0, beta_start: float = 1e-4,
             beta_end: float = 2e-2) -> Dict[str, Tensor]:
    """forward.py:15-27 (fp32, same op order)."""
    beta = torch.linspace(scale * beta_start, scale * beta_end, timesteps, dtype=torch.float32)
    alpha = 1 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return {
        "beta": beta,
        "alpha": alpha,
        "alpha_bar": alpha_bar,
        "sqrt_alpha_bar": torch.sqrt(alpha_bar),
        "one_by_sqrt_alpha": 1.0 / torch.sqrt(alpha),
        "sqrt_one_minus_alpha_bar": torch.sqrt(1 - alpha_bar),
    }


def q_sample(s: Dict[str, Tensor], x0: Tensor, t: Tensor, eps: Tensor) -> Tensor:
    """forward.py:29-37 with the noise passed in."""
    mean = s["sqrt_alpha_bar"].gather(-1, t).reshape(-1, 1, 1, 1, 1) * x0
    std = s["sqrt_one_minus_alpha_bar"].gather(-1, t).reshape(-1, 1, 1, 1, 1)
    return mean + std * eps


def ddpm_step(s: Dict[str, Tensor], eps_pred: Tensor, x: Tensor, t: int, z: Optional[Tensor]) -> Tensor:
    """ddpm.py:25-38 (z = None means the t == 0 zeros)."""
    beta_t = s["beta"][t].reshape(-1, 1, 1, 1, 1)
    a = s["one_by_sqrt_alpha"][t].reshape(-1, 1, 1, 1, 1)
    b = s["sqrt_one_minus_alpha_bar"][t].reshape(-1, 1, 1, 1, 1)
    zz = torch.zeros_like(x) if z is None else z
    return a * (x - (beta_t / b) * eps_pred) + torch.sqrt(beta_t) * zz


Denoiser = Callable[[Tensor, Tensor, Tensor], Tensor]   # (x, t[B] int64, past) -> eps


def generate_ddpm(denoiser: Denoiser, s: Dict[str, Tensor], past: Tensor, x_T: Tensor,
                  noise: Sequence[Tensor], guidance: str = "None", lam: float = 0.0,
                  history: bool = False):
    """ddpm.py:206-236.  noise[i] is the z drawn at loop iteration i (t = T-1-i); the last
    iteration (t = 0) uses none."""
    T = s["beta"].numel()
    x = x_T
    hist = [x]
    for i, t in enumerate(reversed(range(T))):
        tt = torch.full((x.shape[0],), t, dtype=torch.long)
        eps = denoiser(x, tt, past)
        x = ddpm_step(s, eps, x, t, noise[i] if t > 0 else None)
        if guidance == "Sparsity":                        # ddpm.py:223-226, guidance.py:4-8
            sigma = torch.sqrt(s["beta"][t])
            g = torch.zeros_like(x)
            g[:, 0] = torch.sign(x[:, 0])
            x = x - lam * sigma * g
        if history:
            hist.append(x)
    if not history:
        hist.append(x)
    return x, hist


def generate_ddim(denoiser: Denoiser, s: Dict[str, Tensor], past: Tensor, x_T: Tensor,
                  noise: Sequence[Tensor], taus: np.ndarray, sigma_t: float,
                  guidance: str = "None", lam: float = 0.0):
    """ddpm.py:238-282 (DDIM eq. 12; coefficients start at T-1, then follow reversed(taus))."""
    T = s["beta"].numel()
    x = x_T
    beta_t = s["beta"][T - 1]
    sab_t = s["sqrt_alpha_bar"][T - 1]
    somab_t = s["sqrt_one_minus_alpha_bar"][T - 1]
    for i, t in enumerate(reversed(list(taus))):
        t = int(t)
        tt = torch.full((x.shape[0],), t, dtype=torch.long)
        eps = denoiser(x, tt, past)
        beta_prev = s["beta"][t]
        sab_prev = s["sqrt_alpha_bar"][t]
        somab_prev = s["sqrt_one_minus_alpha_bar"][t]
        x0 = (x - somab_t * eps) / sab_t
        direction = torch.sqrt(1 - sab_prev ** 2 - sigma_t ** 2) * eps
        x = sab_prev * x0 + direction + sigma_t * noise[i]
        if guidance == "Sparsity":
            g = torch.zeros_like(x)
            g[:, 0] = torch.sign(x[:, 0])
            x = x - lam * torch.sqrt(beta_t) * g
        beta_t, sab_t, somab_t = beta_prev, sab_prev, somab_prev
    return x


def train_loss(denoiser: Denoiser, s: Dict[str, Tensor], future: Tensor, past: Tensor, t: Tensor,
               eps: Tensor) -> Tensor:
    """ddpm.py:111-121 with t and eps injected (autocast is a no-op on CPU)."""
    x_t = q_sample(s, future, t, eps)
    return F.mse_loss(denoiser(x_t, t, past), eps)


def synthetic_macroprops(n: int, channels: int, rows: int, cols: int, frames: int, seed: int) -> Tensor:
    """SURVEY.md §8(d) synthetic inputs: occupancy Bernoulli(0.2); rho = mask*(1+Poisson(0.5));
    vx, vy = mask*N(0, 0.5^2).  Layout [n, C, H, W, L] fp32."""
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand(n, 1, rows, cols, frames, generator=g) < 0.2).float()
    rho = mask * (1 + torch.poisson(torch.full_like(mask, 0.5), generator=g))
    v = mask * torch.randn(n, max(channels - 1, 1), rows, cols, frames, generator=g) * 0.5
    return torch.cat([rho, v], dim=1)[:, :channels].contiguous()
