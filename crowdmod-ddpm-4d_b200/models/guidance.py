"""Sampling guidance — the part of reference models/guidance.py on the hot path.

``sparsityGradient`` (guidance.py:4-8) is kept for API compatibility; inside the fused chain it
is one extra term of the reverse-step epilogue (coefficient ``g`` of cm_chain_args.coef).
``mass_preservation`` (guidance.py:44-70, a 3 888-iteration python finite-difference loop per
step) is out of scope and raises.
"""
import torch


def sparsityGradient(xnoisy: torch.Tensor, cfg=None, device=None) -> torch.Tensor:
    grad = torch.zeros_like(xnoisy)
    grad[:, 0] = torch.sign(xnoisy[:, 0])
    return grad


def preservationMassNumericalGradientOptimal(*args, **kwargs):
    raise NotImplementedError(
        "mass_preservation guidance is outside the B200 hot path (SURVEY.md §2 #6); use the "
        "reference implementation for it")
