"""Shared helpers for the parity tests (golden loading, seeded weights, reference noise order)."""
import hashlib
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    arrays = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if k != "meta"}
    return meta, arrays


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_unet(meta):
    """Product UNet initialised under the golden's seed; the SHA-256 check proves its weights
    are the ones the reference had when the golden vector was generated."""
    from crowdmod_ddpm_4d_b200.models.backbones.unet import UNet
    torch.manual_seed(meta["seed"])
    net = UNet(**meta["unet"])
    assert sd_hash(net.state_dict()) == meta["sd_sha256"], "seeded init differs from the reference's"
    return net


def structure(meta):
    return dict(num_res_blocks=meta["unet"]["num_res_blocks"],
                num_levels=len(meta["unet"]["base_channels_multiples"]))


def chain_noise(meta):
    """x_T and per-step z in the reference's draw order."""
    shape = (meta["n"], 3, meta["rows"], meta["cols"], meta["F"])
    torch.manual_seed(meta["noise_seed"])
    x_T = torch.randn(shape)
    if meta["sampler"] == "DDPM":
        zs = [torch.randn(shape) for _ in range(meta["T"] - 1)]
    else:
        ntaus = len(np.arange(0, meta["T"] - 1, meta["divider"]))
        zs = [torch.randn(shape) for _ in range(ntaus)]
    return x_T, zs


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def build_dit(meta):
    """Product DiT4D_V4 initialised under the golden's seed (SHA-256 of the seeded init == the reference's), then the
    zero-initialised AdaLN / final projections filled exactly as oracle/make_golden.py::dit_case did."""
    from crowdmod_ddpm_4d_b200.models.backbones.DiT4D_V4 import DiT4D_V4
    from oracle import dit_oracle as dto
    torch.manual_seed(meta["seed"])
    net = DiT4D_V4(**meta["kw"])
    assert len(net.state_dict()) == meta["n_keys"]
    assert sd_hash(net.state_dict()) == meta["init_sha256"], "seeded DiT init differs from the reference's"
    sd = dto.randomize_zero_init({k: v.detach().clone() for k, v in net.state_dict().items()}, meta["seed"])
    net.load_state_dict(sd)
    return net, sd
