"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one rank per GPU, NCCL).

SURVEY.md §4 / §8(e):
  (a) training: the data-parallel gradient (global batch split over the ranks, ONE all-reduce of the flat
      gradient buffer) equals the single-GPU gradient of the whole batch;
  (b) sampling: N shards with global Philox offsets equal the single-GPU chain of the whole sample batch.
Exits non-zero on any mismatch; rank 0 prints the measured differences.
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ddpm_oracle as do  # noqa: E402
from tests._util import build_unet, load_golden, rel_l2  # noqa: E402


def flat_grads(net):
    return torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None]).clone()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from crowdmod_ddpm_4d_b200 import _native as nat
    from crowdmod_ddpm_4d_b200.models.backbones.unet_autograd import disable_data_parallel, enable_data_parallel
    from crowdmod_ddpm_4d_b200.models.diffusion.ddpm import DDPM, ddpm_coefficients, shard_samples

    meta, _ = load_golden("train_atc_b2")
    net = build_unet(meta).to(dev).train()          # same seed on every rank -> identical replicas
    net.dropout_rate = 0.0
    B = 2 * world
    s = do.schedule(1000, 0.5)
    fut = do.synthetic_macroprops(B, 3, 12, 36, 3, 501)
    past = do.synthetic_macroprops(B, 3, 12, 36, 5, 502)
    g = torch.Generator().manual_seed(503)
    t = torch.randint(0, 1000, (B,), generator=g)
    eps = torch.randn(fut.shape, generator=g)
    x_t = do.q_sample(s, fut, t, eps)

    def grads_of(lo, hi):
        net.zero_grad(set_to_none=True)
        loss = F.mse_loss(net(x_t[lo:hi].to(dev), t[lo:hi].to(dev), past[lo:hi].to(dev)), eps[lo:hi].to(dev))
        loss.backward()
        return flat_grads(net), loss.detach()

    g_full, l_full = grads_of(0, B)                  # single-GPU reference: the whole batch, no collective
    enable_data_parallel(net)
    lo, hi = shard_samples(B, rank, world)
    g_dp, l_dp = grads_of(lo, hi)
    disable_data_parallel(net)
    dist.all_reduce(l_dp)
    e_grad = rel_l2(g_dp, g_full)
    e_loss = abs(l_dp.item() / world - l_full.item()) / abs(l_full.item())
    # every rank must hold the same averaged gradient
    g0 = g_dp.clone()
    dist.broadcast(g0, src=0)
    same = torch.equal(g0, g_dp)

    # ---- sharded sampling vs the whole batch on one GPU ----
    n, T = 3 * world, 8
    net.eval()
    ts, coef = ddpm_coefficients(DDPM(timesteps=T, scale=0.5))
    gx = torch.Generator().manual_seed(504)
    xT = torch.randn(n, 3, 12, 36, 3, generator=gx)
    pp = do.synthetic_macroprops(n, 3, 12, 36, 5, 505)
    x_all = xT.to(dev).contiguous()
    net.sample_chain(pp.to(dev).contiguous(), x_all, ts, coef, mode=0, seed=99, sample_offset=0)
    lo, hi = shard_samples(n, rank, world)
    x_sh = xT[lo:hi].to(dev).contiguous()
    net.sample_chain(pp[lo:hi].to(dev).contiguous(), x_sh, ts, coef, mode=0, seed=99, sample_offset=lo)
    parts = [torch.empty_like(x_sh) for _ in range(world)]
    dist.all_gather(parts, x_sh)
    e_chain = rel_l2(torch.cat(parts), x_all)
    dev_err = nat.lib().cm_device_error()
    if rank == 0:
        print(f"MULTIRANK world={world}: DP grad vs single-GPU rel-L2 {e_grad:.3e}, loss rel {e_loss:.3e}, "
              f"replicas identical {same}, sharded chain vs whole batch rel-L2 {e_chain:.3e}", flush=True)
    ok = e_grad <= 5e-5 and e_loss <= 1e-6 and same and e_chain <= 1e-4 and dev_err == 0
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
