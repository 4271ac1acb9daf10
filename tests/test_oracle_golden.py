"""CPU: the oracle restatement (oracle/*.py) against the golden vectors generated from the
unmodified reference (oracle/make_golden.py).  Tolerance: fp32 round-off of two equivalent CPU
evaluations (rel-L2 <= 5e-6; measured ~9e-7)."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as do
from oracle import unet_oracle as uo
from tests._util import build_unet, chain_noise, load_golden, rel_l2, structure


@pytest.mark.parametrize("name", ["unet_atc_b2", "unet_small_b3", "unet_hermes_b1"])
def test_unet_forward_matches_reference_golden(name):
    meta, a = load_golden(name)
    sd = build_unet(meta).state_dict()
    with torch.no_grad():
        eps = uo.unet_forward(sd, a["future"], a["t"], a["past"], **structure(meta))
    assert eps.shape == a["eps"].shape
    assert rel_l2(eps, a["eps"]) <= 5e-6


@pytest.mark.parametrize("name", ["chain_small_ddpm", "chain_small_sparsity", "chain_small_ddim", "chain_atc_T16"])
def test_chain_matches_reference_golden(name):
    meta, a = load_golden(name)
    sd = build_unet(meta).state_dict()
    s = do.schedule(meta["T"], meta["scale"])
    x_T, zs = chain_noise(meta)
    den = lambda x, t, p: uo.unet_forward(sd, x, t, p, **structure(meta))
    with torch.no_grad():
        if meta["sampler"] == "DDPM":
            x0, _ = do.generate_ddpm(den, s, a["past"], x_T, zs, meta["guidance"], meta["lambda"])
        else:
            taus = np.arange(0, meta["T"] - 1, meta["divider"])
            x0 = do.generate_ddim(den, s, a["past"], x_T, zs, taus, meta["sigma"], meta["guidance"], meta["lambda"])
    assert rel_l2(x0, a["x0"]) <= 2e-5


@pytest.mark.parametrize("name", ["train_small", "train_atc_b2"])
def test_train_step_matches_reference_golden(name):
    meta, a = load_golden(name)
    net = build_unet(meta)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "time_blocks.0" not in k)
          for k, v in net.state_dict().items()}
    s = do.schedule(meta["T"], meta["scale"])
    torch.manual_seed(meta["rng_seed"])
    t = torch.randint(0, meta["T"], (meta["B"],))
    eps = torch.randn_like(a["future"])
    den = lambda x, tt, p: uo.unet_forward(sd, x, tt, p, **structure(meta))
    loss = do.train_loss(den, s, a["future"], a["past"], t, eps)
    loss.backward()
    assert abs(loss.item() - float(a["loss"])) <= 1e-5 * abs(float(a["loss"]))
    g = torch.Generator().manual_seed(meta["proj_seed"])
    for i, k in enumerate(meta["names"]):
        grad = sd[k].grad
        r = torch.randn(grad.shape, generator=g)
        n_ref, p_ref = float(a["grad_norms"][i]), float(a["grad_proj"][i])
        assert abs(grad.double().norm().item() - n_ref) <= 1e-4 * n_ref + 1e-9, k
        assert abs((grad.double() * r.double()).sum().item() - p_ref) <= 2e-4 * n_ref * r.norm().item() + 1e-9, k


def test_schedule_matches_reference_golden():
    meta, a = load_golden("schedule_T1000_s0p5")
    s = do.schedule(meta["T"], meta["scale"])
    idx = a["idx"].long()
    for k in ("beta", "alpha", "alpha_bar", "sqrt_alpha_bar", "one_by_sqrt_alpha", "sqrt_one_minus_alpha_bar"):
        assert torch.equal(s[k][idx], a[k]), k
    # anchors quoted in SURVEY.md §8c
    assert abs(s["beta"][0].item() - 5e-5) < 1e-9 and abs(s["beta"][999].item() - 1e-2) < 1e-8
    assert abs(s["alpha_bar"][999].item() - 6.4618e-3) < 1e-6


def test_metrics_oracle_vs_reference_golden():
    """oracle/metrics_oracle.py against the reference's MetricsGenerator outputs (metricsGenerator.py:120-186,
    293-339) on the seeded pair, including the empty-mask frame whose masked PSNR is nan in the reference."""
    import json
    import os
    import numpy as np
    from oracle import metrics_oracle as mo
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_small.npz"))
    meta = json.loads(bytes(g["meta"]).decode())
    pred, gt = mo.synthetic_pair(meta["n"], meta["rows"], meta["cols"], meta["F"], meta["seed"])
    assert np.allclose(mo.mprops_ranges(gt), meta["ranges"], rtol=0, atol=0)
    with np.errstate(all="ignore"):
        a = mo.compute_psnr_metric(pred, gt, meta["chunk"], meta["eps"])
        b = mo.compute_psnr_metric(pred, gt, meta["chunk"], meta["eps"], masked=True)
    for k, v in zip(["PSNR", "MAX_PSNR", "PSNR_OVER_TIME", "MAX_PSNR_OVER_TIME"], a):
        np.testing.assert_allclose(v, g[k], rtol=0, atol=1e-12)
    for k, v in zip(["MASK_PSNR", "MAX_MASK_PSNR", "MASK_PSNR_OVER_TIME", "MAX_MASK_PSNR_OVER_TIME"], b):
        assert np.isnan(g[k]).sum() > 0                      # the empty frame
        np.testing.assert_allclose(v, g[k], rtol=0, atol=1e-12, equal_nan=True)
    re, mre = mo.compute_re_density(pred, gt, meta["chunk"], meta["eps"])
    np.testing.assert_array_equal(re, g["RE_DENSITY"])
    np.testing.assert_array_equal(mre, g["MIN_RE_DENSITY"])
    np.testing.assert_array_equal(mo.compute_tv_metric(pred, gt), g["TV_OVER_TIME"])


@pytest.mark.parametrize("name", ["dit_atc_b2", "dit_small_b3"])
def test_dit_oracle_and_mirror_module_vs_reference_golden(name):
    """oracle/dit_oracle.py reproduces the reference DiT4D_V4 forward (DiT4D_V4.py:348-375) and the product's mirror
    module has the reference's state_dict keys and seeded initialisation bit for bit."""
    from oracle import dit_oracle as dto
    from tests._util import build_dit, load_golden
    meta, a = load_golden(name)
    net, sd = build_dit(meta)
    kw = meta["kw"]
    with torch.no_grad():
        eps = dto.dit_forward(sd, a["future"], torch.tensor(meta["t"]), a["past"], patch=kw["patch_size"],
                              t_patch=kw["t_patch_size"], heads=kw["num_heads"], depth=kw["depth"])
    err = ((eps - a["eps"]).norm() / a["eps"].norm()).item()
    assert err <= 2e-6, err
    with pytest.raises(RuntimeError):
        net(a["future"], torch.tensor(meta["t"]), a["past"])          # CPU tensors: no CPU path
