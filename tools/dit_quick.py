import sys, time, torch
sys.path.insert(0, "."); 
from bench import synthetic_macroprops, WORKLOADS, T_STEPS, SCALE
from crowdmod_ddpm_4d_b200.models.backbones.DiT4D_V4 import DiT4D_V4
from crowdmod_ddpm_4d_b200.models.diffusion.ddpm import DDPM, ddpm_coefficients
dev = torch.device("cuda", 0)
w = WORKLOADS["atc"]
kw = dict(input_channels=3, output_channels=3, grid_rows=12, grid_cols=36, past_len=5, future_len=3, t_patch_size=4, patch_size=4, hidden_size=256, depth=6, num_heads=4, mlp_ratio=4.0, dropout_rate=0.1, time_multiple=4)
torch.manual_seed(42)
net = DiT4D_V4(**kw).to(dev).eval()
ts, coef = ddpm_coefficients(DDPM(timesteps=T_STEPS, scale=SCALE))
for n in (64, 1280):
    past = synthetic_macroprops(n, 3, 12, 36, 5, 77, dev); x = torch.randn(n, 3, 12, 36, 3, device=dev)
    net.sample_chain(past, x, ts[:8], coef[:8]); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 200
    a.record(); net.sample_chain(past, x, ts[:steps], coef[:steps]); b.record(); torch.cuda.synchronize()
    print(f"n={n}: {a.elapsed_time(b)/steps:.4f} ms/step -> {n/(a.elapsed_time(b)/steps*1e-3*T_STEPS):.1f} seq/s")
