"""GPU-resident macro-property dataset and window loader.

Mirrors the reference's data feed for the training / sampling loops -- `MacropropsDataset`
(/root/reference/utils/dataset.py:22-53) wrapped in a `DataLoader(dataset, batch_size, **cfg.DATASET.params)`
(:169-190) whose batches `_train_one_epoch` copies to the device every step
(/root/reference/models/diffusion/ddpm.py:136-137) -- with the raw sequences resident in HBM: a batch is ONE
`cm_window_gather` launch, no worker processes, no host->device copy.  Same windows, same batch order (the sampler
consumes the torch RNG exactly as `RandomSampler` does), bit-identical tensors, already on the device, so
`DDPM_model.train(loader)` / `generate_metrics(loader, ...)` take it unchanged.

There is no CPU fallback: the gather is the native kernel and the tensors live on a CUDA device.
"""
from __future__ import annotations

import torch

from .. import _native


class GpuMacropropsDataset:
    """Same constructor arguments, `indices`, `__len__` and `__getitem__` result as the reference's
    `MacropropsDataset` (dataset.py:22-53); `seq_all` is moved to `device` once."""

    def __init__(self, seq_all, cfg, mprops_count, stride=10, device="cuda"):
        self.mprops_count = mprops_count
        self.stride = stride
        self.past_len = cfg.DATASET.PAST_LEN
        self.future_len = cfg.DATASET.FUTURE_LEN
        seq = torch.as_tensor(seq_all)
        if seq.dim() != 5:
            raise ValueError(f"seq_all must be [N, C, ROWS, COLS, RAW_SEQ_LEN], got {tuple(seq.shape)}")
        total_len = seq.shape[-1]
        window_len = self.past_len + self.future_len
        self.indices = [(s, t) for s in range(seq.shape[0]) for t in range(0, total_len - window_len + 1, stride)]
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _native.NativeError("GpuMacropropsDataset needs a CUDA device (the window gather has no CPU path)")
        self.seq_all = seq.to(device=dev, dtype=torch.float32).contiguous()
        idx = torch.tensor(self.indices, dtype=torch.int32).reshape(-1, 2)
        self._seq_idx = idx[:, 0].contiguous().to(dev)
        self._t0 = idx[:, 1].contiguous().to(dev)

    def __len__(self):
        return len(self.indices)

    def gather(self, sample_ids):
        """(past [b, C, R, Cc, P], future [b, C, R, Cc, F]) of the windows `sample_ids` (device int64 / list)."""
        ids = torch.as_tensor(sample_ids, device=self.seq_all.device).long().reshape(-1).contiguous()
        b = ids.numel()
        n, c, r, cc, t = self.seq_all.shape
        past = torch.empty(b, c, r, cc, self.past_len, device=self.seq_all.device)
        future = torch.empty(b, c, r, cc, self.future_len, device=self.seq_all.device)
        _native.check(_native.lib().cm_window_gather(_native.ptr(self.seq_all), n, c, r, cc, t, _native.ptr(self._seq_idx),
                                                     _native.ptr(self._t0), _native.ptr(ids), b, self.past_len,
                                                     self.future_len, _native.ptr(past), _native.ptr(future),
                                                     _native.current_stream()))
        return past, future

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        past, future = self.gather([idx])
        return past[0], future[0]


class GpuWindowLoader:
    """Iterable of (past, future) device batches in the order `DataLoader(dataset, batch_size, shuffle=...,
    drop_last=...)` yields them (RandomSampler draws its seed from the global torch RNG, then torch.randperm)."""

    def __init__(self, dataset: GpuMacropropsDataset, batch_size, shuffle=False, drop_last=False, generator=None,
                 **_ignored):   # num_workers / pin_memory / ... of cfg.DATASET.params have no meaning here
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.drop_last = bool(drop_last)
        self.generator = generator

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _order(self):
        n = len(self.dataset)
        if not self.shuffle:
            torch.empty((), dtype=torch.int64).random_()          # the base-seed draw happens without shuffling too
            return torch.arange(n)
        g = self.generator
        if g is None:
            # consume the global RNG exactly as a DataLoader epoch does: the iterator's worker base seed first
            # (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__), then RandomSampler's own seed
            torch.empty((), dtype=torch.int64).random_()
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
        return torch.randperm(n, generator=g)

    def __iter__(self):
        order = self._order().to(self.dataset.seq_all.device)
        for b0 in range(0, order.numel(), self.batch_size):
            sel = order[b0:b0 + self.batch_size]
            if self.drop_last and sel.numel() < self.batch_size:
                return
            yield self.dataset.gather(sel)


def gpu_resident(loader_or_dataset, cfg=None, batch_size=None, device="cuda", **params):
    """Drop-in conversion of what the reference's `getDataset` returns: a `DataLoader` over a `MacropropsDataset`
    (or the dataset itself) becomes a `GpuWindowLoader` with the same windows, batch size, shuffle and drop_last."""
    ds = getattr(loader_or_dataset, "dataset", loader_or_dataset)

    class _Cfg:   # the two fields the dataset constructor reads
        class DATASET:
            PAST_LEN = ds.past_len
            FUTURE_LEN = ds.future_len

    gds = GpuMacropropsDataset(ds.seq_all, cfg or _Cfg, ds.mprops_count, stride=ds.stride, device=device)
    assert gds.indices == [tuple(i) for i in ds.indices]
    if hasattr(loader_or_dataset, "dataset"):
        ld = loader_or_dataset
        shuffle = type(getattr(ld, "sampler", None)).__name__ == "RandomSampler"
        params.setdefault("shuffle", shuffle)
        params.setdefault("drop_last", bool(getattr(ld, "drop_last", False)))
        batch_size = batch_size or ld.batch_size
    return GpuWindowLoader(gds, batch_size or len(gds), **params)
