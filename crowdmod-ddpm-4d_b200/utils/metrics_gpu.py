"""On-GPU metrics tail of `generate_metrics`.

Mirrors the reduction metrics of the reference's `MetricsGenerator`
(/root/reference/utils/metrics/metricsGenerator.py): macro-property ranges (:43-68), PSNR / MASK_PSNR with their
MAX_* and *_OVER_TIME variants (:70-86, :120-186), RE_DENSITY / MIN_RE_DENSITY (:293-319) and TV_OVER_TIME
(:88-92, :321-339) -- same method names, same `data_dict` keys, array shapes and column order, same CSV files.
The reference walks the samples calling `.cpu().numpy()` and numpy reductions per frame; here predictions and
ground truth stay on the device as two [n, C, ROWS, COLS, F] tensors, ONE `cm_metrics_reduce` launch produces the
per-(sample, frame) sums in fp64 and the closed forms are evaluated on the [n, F, 21] result.

SSIM (scikit-image windows), the motion-feature histograms and the energy metric are not reductions of this kind:
they are delegated to the reference's own class when it is importable.
"""
from __future__ import annotations

import json
import logging
import os

import numpy as np
import torch

from .. import _native


class GpuMetricsGenerator:
    HEADERS = {
        "PSNR": "rho,vx,vy",
        "MASK_PSNR": "rho,vx,vy",
        "MAX_PSNR": "rho,vx,vy",
        "MAX_MASK_PSNR": "rho,vx,vy",
        "RE_DENSITY": "re_f6,re_f7,re_f8",
        "MIN_RE_DENSITY": "re_f6,re_f7,re_f8",
        "PSNR_OVER_TIME": "rho_f6,vx_f6,vy_f6,rho_f7,vx_f7,vy_f7,rho_f8,vx_f8,vy_f8",
        "MASK_PSNR_OVER_TIME": "rho_f6,vx_f6,vy_f6,rho_f7,vx_f7,vy_f7,rho_f8,vx_f8,vy_f8",
        "TV_OVER_TIME": "rho_f6,vx_f6,vy_f6,rho_f7,vx_f7,vy_f7,rho_f8,vx_f8,vy_f8",
        "MAX_PSNR_OVER_TIME": "rho_f6,vx_f6,vy_f6,rho_f7,vx_f7,vy_f7,rho_f8,vx_f8,vy_f8",
        "MAX_MASK_PSNR_OVER_TIME": "rho_f6,vx_f6,vy_f6,rho_f7,vx_f7,vy_f7,rho_f8,vx_f8,vy_f8",
    }
    GPU_METRICS = ("PSNR", "MASK_PSNR", "RE_DENSITY", "TV")

    def __init__(self, pred_seq_list, gt_seq_list, metrics_params, output_dir=None):
        """`pred_seq_list` / `gt_seq_list`: lists of per-sample [C, ROWS, COLS, F] device tensors (what the reference
        passes) or already stacked [n, C, ROWS, COLS, F] device tensors."""
        self.pred = self._stack(pred_seq_list)
        self.gt = self._stack(gt_seq_list)
        if self.pred.shape != self.gt.shape:
            raise ValueError(f"prediction {tuple(self.pred.shape)} and ground truth {tuple(self.gt.shape)} differ")
        if not self.pred.is_cuda:
            raise _native.NativeError("GpuMetricsGenerator needs CUDA tensors (there is no CPU path)")
        self.params = metrics_params
        self.output_dir = output_dir
        self.data_dict = {name: None for name in self.HEADERS}
        self.mprops_count = int(getattr(metrics_params, "MPROPS_COUNT", 3))
        self._sums = self._reduce()                    # [n, F, 21] float64, host
        mn = self._sums[:, :, 15::2].min(axis=(0, 1))
        mx = self._sums[:, :, 16::2].max(axis=(0, 1))
        self.rho_range, self.vx_range, self.vy_range = (float(mx[c] - mn[c]) for c in range(3))

    @staticmethod
    def _stack(x):
        t = torch.stack(list(x)) if isinstance(x, (list, tuple)) else x
        return t.detach().float().contiguous()

    def _reduce(self):
        n, c, rows, cols, frames = self.pred.shape
        out = torch.empty(n, frames, _native.CM_METRICS_PER_FRAME, dtype=torch.float64, device=self.pred.device)
        _native.check(_native.lib().cm_metrics_reduce(_native.ptr(self.pred), _native.ptr(self.gt), n, c, rows, cols,
                                                      frames, _native.ptr(out), _native.current_stream()))
        return out.cpu().numpy()                        # the one device->host read of the metrics tail (n*F*21 doubles)

    # ----------------------------- metrics (same names / results as the reference) -----------------------------
    @staticmethod
    def _chunk(a, chunk, fn):
        n = a.shape[0]
        out = np.zeros((n // chunk, a.shape[1]))
        for i in range(0, n - n % chunk, chunk):
            out[i // chunk] = fn(a[i:i + chunk], axis=0)
        return out

    def compute_psnr_metric(self, chunkRepdPastSeq, eps, masked_flag=False):
        n, _, rows, cols, F = self.pred.shape
        s = self._sums
        ranges = np.array([self.rho_range, self.vx_range, self.vy_range])
        with np.errstate(invalid="ignore", divide="ignore"):
            mse = s[:, :, 3:6] / s[:, :, 6:7] if masked_flag else s[:, :, 0:3] / float(rows * cols)
            err = np.maximum(mse, eps)                  # an empty mask gives nan, as the reference's max(nan, eps)
            frame = 20 * np.log10(ranges)[None, None, :] - 10 * np.log10(err)      # [n, F, 3]
        over_time = np.zeros((n, self.mprops_count * F))
        for j in range(F):
            over_time[:, j * self.mprops_count:j * self.mprops_count + 3] = frame[:, j, :]
        per = np.zeros((n, self.mprops_count))
        acc = np.zeros((n, 3))
        for j in range(F):                              # the reference accumulates frame by frame
            acc = acc + frame[:, j, :]
        per[:, :3] = acc / F
        logging.info(f'Range of macroprops \n rho:{self.rho_range:.4f}, vx:{self.vx_range:.4f} and vy:{self.vy_range:.4f}')
        pre = "MASK_" if masked_flag else ""
        self.data_dict[pre + "PSNR"] = per
        self.data_dict["MAX_" + pre + "PSNR"] = self._chunk(per, chunkRepdPastSeq, np.max)
        self.data_dict[pre + "PSNR_OVER_TIME"] = over_time
        self.data_dict["MAX_" + pre + "PSNR_OVER_TIME"] = self._chunk(over_time, chunkRepdPastSeq, np.max)

    def compute_re_density_metric(self, chunkRepdPastSeq, eps):
        s = self._sums
        re = np.abs(s[:, :, 13] - s[:, :, 14]) / (s[:, :, 14] + eps)
        self.data_dict["RE_DENSITY"] = re
        self.data_dict["MIN_RE_DENSITY"] = self._chunk(re, chunkRepdPastSeq, np.min)

    def compute_tv_metric(self):
        n, _, _, _, F = self.pred.shape
        s = self._sums
        out = np.zeros((n, self.mprops_count * F))
        for j in range(F):
            out[:, j * self.mprops_count:j * self.mprops_count + 3] = np.abs(s[:, j, 7:10] - s[:, j, 10:13])
        self.data_dict["TV_OVER_TIME"] = out

    # ----------------------------- saving (file names / format of metricsGenerator.py:112-115, :343-360) -----------
    def _save_metric_data(self, match, data, metric, header, samples_per_batch):
        tag = match.group() if match is not None else "NA"
        file_name = f"{self.output_dir}/{metric}_NS{samples_per_batch}_{tag}.csv"
        np.savetxt(file_name, data, delimiter=",", header=header, comments="", fmt="%.4f")
        return file_name

    def save_data_metrics(self, match, title, samples_per_batch, extra=None):
        files = {"title": title}
        if extra:
            files.update(extra)
        for name, header in self.HEADERS.items():
            data = self.data_dict[name]
            if data is not None:
                logging.info(f"Saving metric {name}, entries: {data.shape[0]}")
                files[name] = self._save_metric_data(match, data, name, header, samples_per_batch)
        json_path = os.path.join(self.output_dir, "metrics_files.json")
        with open(json_path, "w") as f:
            json.dump(files, f, indent=2)
        logging.info(f"Metrics filenames saved to {json_path}")
        return files


def compute_metrics_gpu(cfg, gen: GpuMetricsGenerator, metric, chunkRepdPastSeq):
    """The GPU part of the reference's `compute_metrics` dispatch (metricsGenerator.py:374-397); returns the metric
    names it covered so that the caller hands the rest (SSIM, motion features, energy) to the reference class."""
    done = []
    if metric in ("PSNR", "ALL"):
        gen.compute_psnr_metric(chunkRepdPastSeq, cfg.MACROPROPS.EPS)
        done.append("PSNR")
    if metric in ("MASK_PSNR", "ALL"):
        gen.compute_psnr_metric(chunkRepdPastSeq, cfg.MACROPROPS.EPS, masked_flag=True)
        done.append("MASK_PSNR")
    if metric in ("RE_DENSITY", "ALL"):
        gen.compute_re_density_metric(chunkRepdPastSeq, cfg.MACROPROPS.EPS)
        done.append("RE_DENSITY")
    if metric in ("TV", "ALL"):
        gen.compute_tv_metric()
        done.append("TV")
    return done
