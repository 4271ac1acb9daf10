"""Multi-process host logic on CPU (gloo, world_size 2), SURVEY.md §8(e):
  * training: the flat gradient buffer is all-reduced to the MEAN over ranks in one call
    (unet_autograd._allreduce_mean), so every rank steps with the global-batch gradient;
  * sampling: contiguous sample shards + global sample offsets cover the batch exactly once.
No CUDA involved: the native forward/backward are covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crowdmod_ddpm_4d_b200.models.backbones.unet_autograd import (_allreduce_mean, disable_data_parallel,
                                                                          enable_data_parallel)
        from crowdmod_ddpm_4d_b200.models.diffusion.ddpm import shard_samples
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(1003, generator=g)
        mine = flat.clone()
        holder = torch.nn.Linear(1, 1)            # stands in for the UNet module carrying the opt-in flag
        _allreduce_mean(holder, flat)             # a process group that merely exists is NOT used
        assert torch.equal(flat, mine)
        enable_data_parallel(holder)
        _allreduce_mean(holder, flat)
        gathered = [torch.zeros(1003) for _ in range(world)]
        dist.all_gather(gathered, mine)
        ref = torch.stack(gathered).mean(0)
        ok_mean = torch.allclose(flat, ref, atol=1e-6)
        lo, hi = shard_samples(13, rank, world)
        spans = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
        dist.all_gather(spans, torch.tensor([lo, hi]))
        cover = sorted((int(a), int(b)) for a, b in spans)
        ok_cover = cover[0][0] == 0 and cover[-1][1] == 13 and all(cover[i][1] == cover[i + 1][0]
                                                                   for i in range(world - 1))
        disable_data_parallel(holder)
        again = mine.clone()
        _allreduce_mean(holder, again)
        out[rank] = (ok_mean and torch.equal(again, mine), ok_cover)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_mean_and_sample_sharding_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] == (True, True) for r in range(world)), dict(out)


def test_allreduce_is_a_noop_without_process_group():
    import pytest
    from crowdmod_ddpm_4d_b200.models.backbones.unet_autograd import _allreduce_mean, enable_data_parallel
    x = torch.arange(5.0)
    holder = torch.nn.Linear(1, 1)
    _allreduce_mean(holder, x)
    assert torch.equal(x, torch.arange(5.0))
    with pytest.raises(RuntimeError):
        enable_data_parallel(holder)             # opt-in needs an initialised process group
