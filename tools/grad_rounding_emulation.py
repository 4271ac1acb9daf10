"""CPU experiment (oracle + operand-rounding emulation): which fp16 rounding dominates the training-gradient error.
Variants: forward activation rounding / wgrad activation operand / wgrad dOut operand (dgrad exact).  Result
(profiles/r2_train_grad_parity.txt): the forward activation rounding alone accounts for all of it.
  python tools/grad_rounding_emulation.py [batch]"""
import sys, torch, torch.nn.functional as F
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as uo, ddpm_oracle as do
from tests._util import build_unet, load_golden, rel_l2, structure
torch.set_num_threads(16)

def r16(x):
    return x.to(torch.float16).to(x.dtype)

class Conv(torch.autograd.Function):
    cfg = dict(fwd=1, wa=1, wd=1)
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        c = Conv.cfg
        xr = r16(x) if c['fwd'] else x
        ctx.save_for_backward(r16(x) if c['wa'] else x, w)
        ctx.stride, ctx.padding = stride, padding
        ctx.xshape = x.shape
        return _orig(xr, w, b, stride=stride, padding=padding)
    @staticmethod
    def backward(ctx, dy):
        xs, w = ctx.saved_tensors
        c = Conv.cfg
        dx = torch.nn.grad.conv3d_input(ctx.xshape, w, dy, stride=ctx.stride, padding=ctx.padding)
        if c['wd']:
            s = 2.0 ** torch.floor(torch.log2(64.0 / dy.abs().max().clamp_min(1e-30)))
            dyr = r16(dy * s) / s
        else:
            dyr = dy
        dw = torch.nn.grad.conv3d_weight(xs, w.shape, dyr, stride=ctx.stride, padding=ctx.padding)
        db = dy.sum(dim=(0, 2, 3, 4))
        return dx, dw, db, None, None

_orig = F.conv3d
def conv3d(x, w, b=None, stride=1, padding=0):
    if w.shape[1] <= 4 or w.shape[0] <= 4:   # first / final conv: fp32 SIMT in the product
        return _orig(x, w, b, stride=stride, padding=padding)
    return Conv.apply(x, w, b, stride, padding)

def grads(meta, a, t, eps, patched, B=None):
    net = build_unet(meta)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "time_blocks.0" not in k)
          for k, v in net.state_dict().items()}
    s = do.schedule(meta["T"], meta["scale"])
    uo.F.conv3d = conv3d if patched else _orig
    den = lambda x, tt, p: uo.unet_forward(sd, x, tt, p, **structure(meta))
    loss = do.train_loss(den, s, a["future"], a["past"], t, eps)
    loss.backward()
    uo.F.conv3d = _orig
    return {k: v.grad for k, v in sd.items() if v.grad is not None}

meta, a = load_golden("train_atc_b2")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if B != 2:
    a = {"future": do.synthetic_macroprops(B, 3, 12, 36, 3, 77), "past": do.synthetic_macroprops(B, 3, 12, 36, 5, 78)}
torch.manual_seed(5)
t = torch.randint(0, 1000, (B,))
eps = torch.randn_like(a["future"])
ref = grads(meta, a, t, eps, False)
for name, cfg in [("current fwd1 wa1 wd1", dict(fwd=1, wa=1, wd=1)), ("precise fwd, wgrad single", dict(fwd=0, wa=1, wd=1)),
                  ("fwd round only", dict(fwd=1, wa=0, wd=0)), ("wgrad act only", dict(fwd=0, wa=1, wd=0)),
                  ("wgrad dout only", dict(fwd=0, wa=0, wd=1))]:
    Conv.cfg = cfg
    g = grads(meta, a, t, eps, True)
    num = sum((g[k].double() - ref[k].double()).pow(2).sum().item() for k in ref)
    den = sum(ref[k].double().pow(2).sum().item() for k in ref)
    errs = sorted(((rel_l2(g[k], ref[k]), k) for k in ref), reverse=True)
    over = sum(1 for e, k in errs if e > 1e-3)
    import statistics
    print(f"B={B} {name:28s} global {((num/den)**0.5):.2e}  worst {errs[0][0]:.2e} ({errs[0][1]})  median {statistics.median(e for e,_ in errs):.2e}  >1e-3: {over}/{len(errs)}")
