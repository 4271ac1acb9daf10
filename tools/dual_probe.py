"""Experiment: one B=64 chain vs two concurrent B=32 chains (two handles, two streams, two host threads)."""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ATC, ROWS, COLS, PAST, FUT, synthetic_macroprops  # noqa: E402
from crowdmod_ddpm_4d_b200.models.backbones.unet import UNet, _NativePlan, _PLANS  # noqa: E402
from crowdmod_ddpm_4d_b200.models.diffusion.forward import ForwardSampler  # noqa: E402
from crowdmod_ddpm_4d_b200.models.diffusion.ddpm import ddpm_coefficients  # noqa: E402

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    parts = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    net = UNet(**ATC).to(dev).eval()
    sampler = ForwardSampler(timesteps=1000, scale=0.5).to(dev)
    tsteps, coef = ddpm_coefficients(sampler)
    tsteps, coef = tsteps[:nsteps], coef[:nsteps]
    past = synthetic_macroprops(n, 3, ROWS, COLS, PAST, 1234, dev)

    def run_single():
        x = torch.randn(n, 3, ROWS, COLS, FUT, device=dev)
        net.sample_chain(past, x, tsteps, coef, mode=0, seed=1, sample_offset=0)
        return x

    key = (ROWS, COLS, PAST, FUT)
    m = n // parts
    plans = [_NativePlan(net, *key) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    main_plan = net._plan(*key)

    import ctypes as C
    from crowdmod_ddpm_4d_b200 import _native as nat
    for pl in plans:
        pl.sync(net, need_table=True)
    ts32 = tsteps.to(torch.int32).contiguous().cpu()
    cf32 = coef.to(torch.float32).contiguous().cpu()

    def chain(pl, ps, xs, off, stream):
        a = nat.ChainArgs()
        a.past = ps.data_ptr(); a.x = xs.data_ptr(); a.n = xs.shape[0]; a.nsteps = ts32.numel()
        a.tsteps = ts32.data_ptr(); a.coef = cf32.data_ptr(); a.mode = 0; a.noise = None
        a.seed = 1; a.sample_offset = off; a.history = None; a.use_graph = 1
        nat.check(nat.lib().cm_ddpm_sample(pl.handle, C.byref(a), C.c_void_p(stream.cuda_stream)))

    def run_multi():
        x = torch.randn(n, 3, ROWS, COLS, FUT, device=dev)
        torch.cuda.synchronize()
        def work(i):
            chain(plans[i], past[i * m:(i + 1) * m], x[i * m:(i + 1) * m], i * m, streams[i])
        ths = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
        for t in ths: t.start()
        for t in ths: t.join()
        torch.cuda.synchronize()
        return x

    outs = {}
    for name, fn in (("single", run_single), ("multi", run_multi), ("single", run_single), ("multi", run_multi)):
        torch.manual_seed(7)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs[name] = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name}: {dt*1e3/nsteps:.4f} ms/step ({n} samples, {nsteps} steps, parts {parts})")
    print("bit-identical:", torch.equal(outs["single"], outs["multi"]))

if __name__ == "__main__":
    main()
