"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reduction metrics of the reference's
`generate_metrics` tail.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this; the product path (crowdmod-ddpm-4d_b200/utils/metrics_gpu.py ->
cm_metrics_reduce) never does.

Follows /root/reference/utils/metrics/metricsGenerator.py:
  * mprops_ranges          <- MetricsGenerator._get_mprops_ranges      (:43-68)
  * psnr / psnr_masked     <- _my_psnr / _my_psnr_masked               (:70-86)
  * compute_psnr_metric    <- compute_psnr_metric                      (:120-186)
  * compute_re_density     <- compute_re_density_metric                (:293-319)
  * total_variation / compute_tv_metric <- _compute_tv / compute_tv_metric (:88-92, :321-339)
Pinned against the live reference (tests/test_oracle_pins.py) and the golden vectors generated from it
(oracle/make_golden.py -> tests/golden/metrics_small.npz).

Inputs are arrays [n, C, ROWS, COLS, F] (float32), C >= 3 = (rho, vx, vy); outputs have the reference's
shapes and column order (rho_f0, vx_f0, vy_f0, rho_f1, ...).
"""
import numpy as np


def mprops_ranges(gt):
    """Global (max - min) of each macro-property over all ground-truth sequences."""
    gt = np.asarray(gt)
    return tuple(float(np.float64(gt[:, c].max()) - np.float64(gt[:, c].min())) for c in range(3))   # fp64 difference of fp32 extrema


def psnr(y_gt, y_hat, data_range, eps):
    err = np.mean((y_gt - y_hat) ** 2, dtype=np.float64)
    err = max(err, eps)
    return 20 * np.log10(data_range) - 10 * np.log10(err)


def psnr_masked(y_gt, y_hat, data_range, eps, mask):
    with np.errstate(invalid="ignore", divide="ignore"):
        err = np.mean((y_gt[mask] - y_hat[mask]) ** 2, dtype=np.float64)   # empty mask -> nan, as the reference
    err = max(err, eps)
    return 20 * np.log10(data_range) - 10 * np.log10(err)


def _chunk_reduce(a, chunk, fn):
    n = a.shape[0]
    out = np.zeros((n // chunk, a.shape[1]))
    for i in range(0, n, chunk):
        if i // chunk < out.shape[0]:
            out[i // chunk] = fn(a[i:i + chunk], axis=0)
    return out


def compute_psnr_metric(pred, gt, chunk, eps, masked=False, mprops_count=3):
    """Returns (PSNR [n, 3], MAX_PSNR [n // chunk, 3], PSNR_OVER_TIME [n, 3F], MAX_PSNR_OVER_TIME)."""
    pred, gt = np.asarray(pred), np.asarray(gt)
    n, F = pred.shape[0], pred.shape[-1]
    ranges = mprops_ranges(gt)
    per = np.zeros((n, mprops_count))
    over_time = np.zeros((n, mprops_count * F))
    for i in range(n):
        acc = np.zeros(3)
        for j in range(F):
            g, p = gt[i, :, :, :, j], pred[i, :, :, :, j]
            mask = g[0] > 0.00001
            for c in range(3):
                v = psnr_masked(g[c], p[c], ranges[c], eps, mask) if masked else psnr(g[c], p[c], ranges[c], eps)
                acc[c] += v
                over_time[i, j * mprops_count + c] = v
        per[i, :3] = acc / F
    return per, _chunk_reduce(per, chunk, np.max), over_time, _chunk_reduce(over_time, chunk, np.max)


def compute_re_density(pred, gt, chunk, eps):
    """Returns (RE_DENSITY [n, F], MIN_RE_DENSITY [n // chunk, F])."""
    pred, gt = np.asarray(pred), np.asarray(gt)
    n, F = pred.shape[0], pred.shape[-1]
    re = np.zeros((n, F))
    for i in range(n):
        p_tot = pred[i, 0].sum(axis=(0, 1))
        g_tot = gt[i, 0].sum(axis=(0, 1))
        re[i] = np.abs(p_tot - g_tot) / (g_tot + eps)
    return re, _chunk_reduce(re, chunk, np.min)


def total_variation(field):
    return np.abs(np.diff(field, axis=0)).sum() + np.abs(np.diff(field, axis=1)).sum()


def compute_tv_metric(pred, gt, mprops_count=3):
    """Returns TV_OVER_TIME [n, 3F]: |TV(pred) - TV(gt)| per (frame, macro-property)."""
    pred, gt = np.asarray(pred), np.asarray(gt)
    n, F = pred.shape[0], pred.shape[-1]
    out = np.zeros((n, mprops_count * F))
    for i in range(n):
        for j in range(F):
            for c in range(mprops_count):
                out[i, j * mprops_count + c] = np.abs(total_variation(pred[i, c, :, :, j]) - total_variation(gt[i, c, :, :, j]))
    return out


def synthetic_pair(n, rows, cols, F, seed, noise=0.3):
    """Seeded (pred, gt) pair of sparse macro-property sequences (same recipe as the bench's synthetic data):
    gt = occupancy mask * (rho >= 1, velocities ~ N(0, 0.5^2)); pred = gt + noise; the first sample's first
    frame is left EMPTY (no occupied cell: the reference's masked PSNR is nan there)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((n, 1, rows, cols, F)) < 0.2
    rho = mask * (1 + rng.poisson(0.5, size=mask.shape))
    vel = mask * rng.normal(0.0, 0.5, size=(n, 2, rows, cols, F))
    gt = np.concatenate([rho, vel], axis=1).astype(np.float32)
    gt[0, :, :, :, 0] = 0.0
    pred = (gt + noise * rng.normal(size=gt.shape)).astype(np.float32)
    return pred, gt
