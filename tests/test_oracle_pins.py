"""CPU, build container only: the oracle restatement against the LIVE reference imported from
/root/reference (skipped where the reference does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ddpm_oracle as do
from oracle import ref_shim
from oracle import unet_oracle as uo

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ns():
    return ref_shim.load()


@pytest.mark.parametrize("kw,geom", [
    (dict(input_channels=3, output_channels=3, num_res_blocks=1, base_channels=32,
          base_channels_multiples=[1, 2, 4], apply_attention=[False, False, True, False],
          dropout_rate=0.1, time_multiple=4, condition="Past"), (8, 12, 5, 3)),          # ETH-UCY shape
    (dict(input_channels=3, output_channels=3, num_res_blocks=2, base_channels=32,
          base_channels_multiples=[1, 2], apply_attention=[True, False],
          dropout_rate=0.1, time_multiple=2, condition="Past"), (4, 6, 3, 1)),
])
def test_unet_forward_vs_live_reference(ns, kw, geom):
    rows, cols, P, F_ = geom
    torch.manual_seed(3)
    ref = ns.UNet(**kw).eval()
    x = torch.randn(2, 3, rows, cols, F_)
    past = do.synthetic_macroprops(2, 3, rows, cols, P, 77)
    t = torch.tensor([5, 900])
    with torch.no_grad():
        a = ref(x, t, past)
        b = uo.unet_forward(ref.state_dict(), x, t, past, num_res_blocks=kw["num_res_blocks"],
                            num_levels=len(kw["base_channels_multiples"]))
    assert uo.rel_l2(b, a) <= 5e-6


def test_product_module_reproduces_reference_init(ns):
    from crowdmod_ddpm_4d_b200.models.backbones.unet import UNet
    kw = dict(input_channels=3, output_channels=3, num_res_blocks=1, base_channels=64,
              base_channels_multiples=[1, 2, 4], apply_attention=[False, False, True, False],
              dropout_rate=0.1, time_multiple=4, condition="Past")
    torch.manual_seed(42)
    a = ns.UNet(**kw).state_dict()
    torch.manual_seed(42)
    b = UNet(**kw).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_schedule_and_step_vs_live_reference(ns):
    ref = ns.DDPM(timesteps=1000, scale=0.5)
    s = do.schedule(1000, 0.5)
    for k, v in s.items():
        assert torch.equal(getattr(ref, k), v), k
    x = torch.randn(2, 3, 4, 4, 2)
    e = torch.randn_like(x)
    torch.manual_seed(9)
    out, _, _ = ref.step(e, x, 321)
    torch.manual_seed(9)
    z = torch.randn_like(x)
    assert torch.equal(out, do.ddpm_step(s, e, x, 321, z))
