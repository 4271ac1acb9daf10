"""Does programmatic dependent launch (CM_PDL=1) change the step time, eagerly and under graph replay?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ATC, ROWS, COLS, PAST, FUT, synthetic_macroprops  # noqa: E402
from crowdmod_ddpm_4d_b200.models.backbones.unet import UNet  # noqa: E402
from crowdmod_ddpm_4d_b200.models.diffusion.forward import ForwardSampler  # noqa: E402
from crowdmod_ddpm_4d_b200.models.diffusion.ddpm import ddpm_coefficients  # noqa: E402
n, nsteps = 64, 300
dev = torch.device("cuda", 0)
torch.manual_seed(42)
net = UNet(**ATC).to(dev).eval()
sampler = ForwardSampler(timesteps=1000, scale=0.5).to(dev)
tsteps, coef = ddpm_coefficients(sampler)
tsteps, coef = tsteps[:nsteps], coef[:nsteps]
past = synthetic_macroprops(n, 3, ROWS, COLS, PAST, 1234, dev)
for use_graph in (True, False):
    for rep in range(2):
        x = torch.randn(n, 3, ROWS, COLS, FUT, device=dev)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        net.sample_chain(past, x, tsteps, coef, mode=0, seed=1, sample_offset=0, use_graph=use_graph)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"CM_PDL={os.environ.get('CM_PDL', '0')} graph={use_graph}: {dt * 1e3 / nsteps:.4f} ms/step", flush=True)
