"""The C-ABI library loads without a GPU and exports every symbol include/ssb.h declares."""
import ctypes
import os.path as osp
import re

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))


def test_libssb_exports_header_symbols():
    import __graft_entry__ as g

    g.build()
    from spark_sched_sim_b200 import _native

    L = _native.lib()
    header = open(osp.join(REPO, "include", "ssb.h")).read()
    declared = set(re.findall(r"\b(ssb_[a-z_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in ssb.h but not exported"
    assert declared == set(_native.EXPORTS)
    assert L.ssb_abi_version() == _native.ABI_VERSION


def test_struct_sizes_match_header():
    from spark_sched_sim_b200 import _native

    assert _native.OBS_HDR_DTYPE.itemsize == 48
    assert _native.STATS_DTYPE.itemsize == 64
    assert ctypes.sizeof(_native.SsbConfig) == 64
    assert ctypes.sizeof(_native.SsbViews) == 56
    assert ctypes.sizeof(_native.SsbDecimaViews) == 64
    assert ctypes.sizeof(_native.SsbPackedObs) == 64


def test_workspace_bytes_no_gpu(bank):
    """ssb_workspace_bytes is pure host arithmetic: callable without a device."""
    from spark_sched_sim_b200 import _native as nat

    cfg = nat.SsbConfig(4096, 10, 50, 50, 0, 0, 2000.0, 1000.0, 4e-5, 0.0)
    bs, keep = nat.make_bank_struct(bank)
    n = ctypes.c_size_t()
    assert nat.lib().ssb_workspace_bytes(ctypes.byref(cfg), ctypes.byref(bs), ctypes.byref(n)) == 0
    assert 100e6 < n.value < 2e9
    dec = nat.SsbConfig(4096, 10, 50, 50, 0, 0, 2000.0, 1000.0, 4e-5, 0.0, nat.FLAG_DECIMA_OBS, 0)
    n2 = ctypes.c_size_t()
    assert nat.lib().ssb_workspace_bytes(ctypes.byref(dec), ctypes.byref(bs), ctypes.byref(n2)) == 0
    assert n2.value > n.value  # the Decima observation buffers are only allocated on request
    bad = nat.SsbConfig(4096, 500, 50, 50, 0, 0, 2000.0, 1000.0, 4e-5, 0.0)
    assert nat.lib().ssb_workspace_bytes(ctypes.byref(bad), ctypes.byref(bs), ctypes.byref(n)) == -1
