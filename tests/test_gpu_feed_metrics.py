"""GPU parity of the two callers either side of the hot path, through the C ABI:
  * cm_window_gather / GpuWindowLoader  vs  oracle/dataset_oracle.py (pinned to the reference's MacropropsDataset +
    DataLoader) and a live torch DataLoader over the same windows: BIT-EXACT batches in the same order;
  * cm_metrics_reduce / GpuMetricsGenerator  vs  oracle/metrics_oracle.py and the golden generated from the reference's
    MetricsGenerator.  Tolerances: PSNR 1e-6 dB absolute (fp64 sums in a different order).  RE_DENSITY and TV are
    differences of sums the reference accumulates in FLOAT32 (numpy .sum() of float32 arrays): its values carry ~1e-6
    (RE) / ~1e-4 (TV) absolute rounding, so the bound is 2e-5 relative + that absolute floor -- far below the %.4f the
    CSV files keep -- and the kernel (fp64 accumulation) is additionally checked against the EXACT fp64 value, where
    it must be at least as close as the reference.  nan positions (empty rho mask) must coincide."""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _cfg(past, future):
    return types.SimpleNamespace(DATASET=types.SimpleNamespace(PAST_LEN=past, FUTURE_LEN=future),
                                 MACROPROPS=types.SimpleNamespace(EPS=1e-6))


class _HostWindows(torch.utils.data.Dataset):
    """The reference's MacropropsDataset restated in the test (utils/dataset.py:22-53) to drive a real DataLoader."""

    def __init__(self, seq, past, future, stride):
        self.seq_all, self.past_len, self.future_len, self.stride, self.mprops_count = seq, past, future, stride, seq.shape[1]
        self.indices = [(s, t) for s in range(seq.shape[0]) for t in range(0, seq.shape[-1] - past - future + 1, stride)]

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        s, t = self.indices[i]
        w = self.seq_all[s, :, :, :, t:t + self.past_len + self.future_len]
        return w[:, :, :, :self.past_len], w[:, :, :, self.past_len:]


@pytest.mark.parametrize("geom", [(3, 3, 4, 5, 23, 5, 3, 4, 4), (5, 3, 12, 36, 40, 5, 3, 8, 64), (2, 4, 8, 12, 17, 8, 8, 1, 3)],
                         ids=["tiny", "atc", "ragged"])
@pytest.mark.parametrize("shuffle,drop_last", [(False, False), (True, True), (True, False)])
def test_window_loader_bit_exact_and_same_order(geom, shuffle, drop_last):
    from crowdmod_ddpm_4d_b200.utils.dataset_gpu import GpuMacropropsDataset, GpuWindowLoader, gpu_resident
    from oracle import dataset_oracle as dso
    n, c, rows, cols, T, P, F, stride, bs = geom
    seq = np.random.default_rng(5).normal(size=(n, c, rows, cols, T)).astype(np.float32)
    ds = GpuMacropropsDataset(seq, _cfg(P, F), c, stride=stride)
    assert ds.indices == dso.window_indices(n, T, P, F, stride)
    torch.manual_seed(77)
    ours = list(GpuWindowLoader(ds, bs, shuffle=shuffle, drop_last=drop_last))
    torch.manual_seed(77)
    ref = list(dso.batches(seq, P, F, stride, bs, shuffle=shuffle, drop_last=drop_last))
    host = _HostWindows(torch.from_numpy(seq), P, F, stride)
    torch.manual_seed(77)
    live = list(torch.utils.data.DataLoader(host, batch_size=bs, shuffle=shuffle, drop_last=drop_last))
    assert len(ours) == len(ref) == len(live) == len(GpuWindowLoader(ds, bs, shuffle=shuffle, drop_last=drop_last))
    for (p0, f0), (p1, f1), (p2, f2) in zip(ours, ref, live):
        assert p0.is_cuda and f0.is_cuda
        assert torch.equal(p0.cpu(), p1) and torch.equal(f0.cpu(), f1)
        assert torch.equal(p0.cpu(), p2) and torch.equal(f0.cpu(), f2)
    # drop-in conversion of an existing DataLoader keeps batch size / shuffle / drop_last
    conv = gpu_resident(torch.utils.data.DataLoader(host, batch_size=bs, shuffle=shuffle, drop_last=drop_last))
    assert (conv.batch_size, conv.shuffle, conv.drop_last) == (bs, shuffle, drop_last)
    p, f = ds[len(ds) - 1]
    assert torch.equal(p.cpu(), host[len(host) - 1][0]) and torch.equal(f.cpu(), host[len(host) - 1][1])


def test_window_gather_rejects_bad_window():
    import crowdmod_ddpm_4d_b200._native as nat
    seq = torch.zeros(1, 3, 2, 2, 6, device="cuda")
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.zeros(1, 3, 2, 2, 5, device="cuda")
    rc = nat.lib().cm_window_gather(nat.ptr(seq), 1, 3, 2, 2, 6, nat.ptr(idx), nat.ptr(idx), None, 1, 5, 3, nat.ptr(out),
                                    nat.ptr(out), nat.current_stream())
    assert rc != 0 and b"does not fit" in nat.lib().cm_last_error()


def _gen(pred, gt):
    from crowdmod_ddpm_4d_b200.utils.metrics_gpu import GpuMetricsGenerator
    return GpuMetricsGenerator(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), types.SimpleNamespace(MPROPS_COUNT=3), None)


def _check(gen, ref, keys_tol):
    for k, (rtol, atol) in keys_tol.items():
        a, b = gen.data_dict[k], ref[k]
        assert a.shape == b.shape, k
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True, err_msg=k)


def test_metrics_vs_reference_golden():
    from oracle import metrics_oracle as mo
    g = np.load(os.path.join(HERE, "golden", "metrics_small.npz"))
    meta = json.loads(bytes(g["meta"]).decode())
    pred, gt = mo.synthetic_pair(meta["n"], meta["rows"], meta["cols"], meta["F"], meta["seed"])
    gen = _gen(pred, gt)
    np.testing.assert_allclose([gen.rho_range, gen.vx_range, gen.vy_range], meta["ranges"], rtol=0, atol=0)
    gen.compute_psnr_metric(meta["chunk"], meta["eps"])
    gen.compute_psnr_metric(meta["chunk"], meta["eps"], masked_flag=True)
    gen.compute_re_density_metric(meta["chunk"], meta["eps"])
    gen.compute_tv_metric()
    psnr = (0, 1e-6)
    _check(gen, g, {"PSNR": psnr, "MAX_PSNR": psnr, "PSNR_OVER_TIME": psnr, "MAX_PSNR_OVER_TIME": psnr, "MASK_PSNR": psnr,
                    "MAX_MASK_PSNR": psnr, "MASK_PSNR_OVER_TIME": psnr, "MAX_MASK_PSNR_OVER_TIME": psnr,
                    "RE_DENSITY": (2e-5, 5e-6), "MIN_RE_DENSITY": (2e-5, 5e-6), "TV_OVER_TIME": (2e-5, 1e-4)})
    assert np.isnan(gen.data_dict["MASK_PSNR_OVER_TIME"]).sum() == 3          # the empty frame of sample 0


@pytest.mark.parametrize("geom", [(6, 8, 12, 3, 3), (64, 12, 36, 3, 4), (5, 28, 24, 8, 5)], ids=["ethucy", "atc", "hermes_f8"])
def test_metrics_vs_oracle(geom):
    from oracle import metrics_oracle as mo
    n, rows, cols, F, chunk = geom
    pred, gt = mo.synthetic_pair(n, rows, cols, F, 1000 + n)
    gen = _gen(pred, gt)
    with np.errstate(all="ignore"):
        a = mo.compute_psnr_metric(pred, gt, chunk, 1e-6)
        b = mo.compute_psnr_metric(pred, gt, chunk, 1e-6, masked=True)
    re, mre = mo.compute_re_density(pred, gt, chunk, 1e-6)
    ref = dict(zip(["PSNR", "MAX_PSNR", "PSNR_OVER_TIME", "MAX_PSNR_OVER_TIME"], a))
    ref.update(zip(["MASK_PSNR", "MAX_MASK_PSNR", "MASK_PSNR_OVER_TIME", "MAX_MASK_PSNR_OVER_TIME"], b))
    ref.update(RE_DENSITY=re, MIN_RE_DENSITY=mre, TV_OVER_TIME=mo.compute_tv_metric(pred, gt))
    gen.compute_psnr_metric(chunk, 1e-6)
    gen.compute_psnr_metric(chunk, 1e-6, masked_flag=True)
    gen.compute_re_density_metric(chunk, 1e-6)
    gen.compute_tv_metric()
    psnr = (0, 1e-6)
    tol = {k: psnr for k in ref if "PSNR" in k}
    tol.update(RE_DENSITY=(2e-5, 5e-6), MIN_RE_DENSITY=(2e-5, 5e-6), TV_OVER_TIME=(2e-5, 1e-4))
    _check(gen, ref, tol)
    # against exact fp64 arithmetic the kernel is at least as accurate as the reference's float32 sums
    p64, g64 = pred.astype(np.float64), gt.astype(np.float64)
    exact = np.abs(p64[:, 0].sum(axis=(1, 2)) - g64[:, 0].sum(axis=(1, 2))) / (g64[:, 0].sum(axis=(1, 2)) + 1e-6)
    err_gpu = np.abs(gen.data_dict["RE_DENSITY"] - exact) / (np.abs(exact) + 1e-6)
    err_ref = np.abs(re - exact) / (np.abs(exact) + 1e-6)
    assert err_gpu.max() <= 1e-12 and err_gpu.max() <= err_ref.max() + 1e-15


def test_metrics_full_size_properties_and_csv(tmp_path):
    """generate_metrics' production batch (n = 1280, ATC shape): identical sequences give the closed-form values
    (PSNR = 20 log10(range) - 10 log10(eps), zero TV / density error), a sample permutation permutes the rows, and the
    CSV files carry the reference's names and headers (metricsGenerator.py:112-115)."""
    import re
    from oracle import metrics_oracle as mo
    from crowdmod_ddpm_4d_b200.utils.metrics_gpu import GpuMetricsGenerator, compute_metrics_gpu
    _, gt = mo.synthetic_pair(1280, 12, 36, 3, 3)
    gen = _gen(gt.copy(), gt)
    gen.compute_psnr_metric(20, 1e-6)
    gen.compute_re_density_metric(20, 1e-6)
    gen.compute_tv_metric()
    want = 20 * np.log10([gen.rho_range, gen.vx_range, gen.vy_range]) - 10 * np.log10(1e-6)
    np.testing.assert_allclose(gen.data_dict["PSNR"], np.broadcast_to(want, (1280, 3)), rtol=0, atol=1e-9)
    assert gen.data_dict["MAX_PSNR"].shape == (64, 3)
    assert not gen.data_dict["TV_OVER_TIME"].any() and not gen.data_dict["RE_DENSITY"].any()
    pred, gt2 = mo.synthetic_pair(256, 12, 36, 3, 4)
    perm = np.random.default_rng(0).permutation(256)
    g0, g1 = _gen(pred, gt2), _gen(pred[perm], gt2[perm])
    cfg = _cfg(5, 3)
    for g in (g0, g1):
        g.output_dir = str(tmp_path)
        assert compute_metrics_gpu(cfg, g, "ALL", 4) == ["PSNR", "MASK_PSNR", "RE_DENSITY", "TV"]
    for k in ("PSNR_OVER_TIME", "MASK_PSNR_OVER_TIME", "RE_DENSITY", "TV_OVER_TIME"):
        assert np.array_equal(g0.data_dict[k][perm], g1.data_dict[k], equal_nan=True), k
    files = g0.save_data_metrics(re.search(r"TE\d+_PL\d+_FL\d+_CE\d+_NA", "x_TE10_PL5_FL3_CE7_NA.pth"), "t", 256)
    assert os.path.basename(files["PSNR"]) == "PSNR_NS256_TE10_PL5_FL3_CE7_NA.csv"
    assert open(files["TV_OVER_TIME"]).readline().strip() == GpuMetricsGenerator.HEADERS["TV_OVER_TIME"]
    assert json.load(open(tmp_path / "metrics_files.json"))["title"] == "t"
