"""CPU: the C-ABI shared library loads without a GPU and exports every symbol that
include/crowdmod_b200.h declares; the ctypes binding covers exactly the same set."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "crowdmod_b200.h")).read()
    return sorted(set(re.findall(r"CM_API\s+[\w\s\*]+?\b(cm_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import crowdmod_ddpm_4d_b200._native as n
    names = declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(n.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(n.SIGNATURES) == names
    assert n.lib().cm_version() >= 100


def test_plan_entry_points_work_without_gpu():
    import crowdmod_ddpm_4d_b200._native as n
    cfg = n.UNetConfig(in_channels=3, out_channels=3, num_res_blocks=1, base_channels=32, num_levels=3,
                       time_multiple=4, rows=12, cols=36, past_len=5, future_len=3, table_steps=1000,
                       weight_terms=2)
    cfg.mult[0], cfg.mult[1], cfg.mult[2] = 1, 2, 4
    cfg.attn[2] = 1
    h = ctypes.c_void_p()
    n.check(n.lib().cm_unet_create(ctypes.byref(cfg), ctypes.byref(h)))
    assert n.lib().cm_unet_param_count(h) == 169
    n.check(n.lib().cm_unet_destroy(h))
    bad = n.UNetConfig(in_channels=3, out_channels=3, num_res_blocks=1, base_channels=24, num_levels=1,
                       time_multiple=4, rows=4, cols=4, past_len=2, future_len=2)
    assert n.lib().cm_unet_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"multiple of 32" in n.lib().cm_last_error()
