"""TEST INFRASTRUCTURE — CPU restatement of the reference's windowed macro-property dataset and of the batch
order its DataLoader produces.  Only tests/, smoke() and bench.py's cpu_baseline
leg may import this; the product path (crowdmod-ddpm-4d_b200/utils/dataset_gpu.py -> cm_window_gather)
never does.

Follows /root/reference/utils/dataset.py:
  * window_indices   <- MacropropsDataset.__init__   (:22-41): (sequence, start) pairs, start stepping by
                        MACROPROPS.STRIDE while a PAST_LEN + FUTURE_LEN window still fits
  * window           <- MacropropsDataset.__getitem__ (:46-53): past = first PAST_LEN frames, future = the rest
  * batches          <- torch.utils.data.DataLoader(dataset, batch_size, shuffle, drop_last) as built at
                        dataset.py:169-190 (`**cfg.DATASET.params`: shuffle / drop_last; the worker count does not
                        change the order), default collate = stack
Pinned against the live reference class and a live DataLoader in tests/test_oracle_pins.py.
"""
import numpy as np
import torch


def window_indices(n_seq, total_len, past_len, future_len, stride):
    win = past_len + future_len
    return [(s, t) for s in range(n_seq) for t in range(0, total_len - win + 1, stride)]


def window(seq_all, idx_pair, past_len, future_len):
    s, t = idx_pair
    w = seq_all[s, :, :, :, t:t + past_len + future_len]
    return w[:, :, :, :past_len], w[:, :, :, past_len:]


def epoch_order(n, shuffle, generator=None):
    """Sample order of one DataLoader epoch (torch.utils.data.RandomSampler / SequentialSampler): with shuffle the
    iterator first draws its worker base seed from the global RNG (torch/utils/data/dataloader.py,
    _BaseDataLoaderIter.__init__), then the sampler draws a seed and yields torch.randperm(n) of a fresh generator."""
    if not shuffle:
        torch.empty((), dtype=torch.int64).random_()          # DataLoader iterator's base-seed draw (no shuffle either)
        return list(range(n))
    if generator is None:
        torch.empty((), dtype=torch.int64).random_()          # the DataLoader iterator's own base-seed draw comes first
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        generator = torch.Generator()
        generator.manual_seed(seed)
    return torch.randperm(n, generator=generator).tolist()


def batches(seq_all, past_len, future_len, stride, batch_size, shuffle=False, drop_last=False):
    """Yields (past [b, C, R, Cc, P], future [b, C, R, Cc, F]) exactly as the reference's DataLoader would."""
    seq_all = np.asarray(seq_all)
    idx = window_indices(seq_all.shape[0], seq_all.shape[-1], past_len, future_len, stride)
    order = epoch_order(len(idx), shuffle)
    for b0 in range(0, len(order), batch_size):
        sel = order[b0:b0 + batch_size]
        if drop_last and len(sel) < batch_size:
            break
        ps, fs = zip(*(window(seq_all, idx[i], past_len, future_len) for i in sel))
        yield torch.from_numpy(np.stack(ps)), torch.from_numpy(np.stack(fs))
