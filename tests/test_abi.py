"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports exactly the symbols
include/discogan_b200.h declares (no compute calls -- there is no GPU here)."""
import re
import subprocess

import pytest

from discogan_modernized_b200 import _lib


@pytest.fixture(scope="module")
def built():
    _lib.build()
    return _lib.lib()


def test_header_parses_every_declaration():
    text = _lib.HEADER.read_text()
    declared = set(re.findall(r"\b(dg_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    protos = _lib.parse_header()
    assert declared == set(protos), declared ^ set(protos)
    assert len(protos) >= 38
    assert protos["dg_last_error"][0].__name__ == "c_char_p"
    assert len(protos["dg_conv4x4s2_wgrad"][1]) == 13      # ..., ws, ws_bytes, opts, stream


def test_library_exports_all_symbols(built):
    for name in _lib.parse_header():
        assert hasattr(built, name), name
    assert built.dg_version() >= 2
    assert built.dg_launch_count() == 0
    # the library records the sources it was compiled from; _lib.lib() refuses a stale one
    assert built.dg_source_hash().decode() == "dgsrc:" + _lib.source_hash() == "dgsrc:" + _lib.built_hash()


def test_no_global_setters_in_the_abi():
    """Round-1's process-global setters are gone: launch options travel with each call (dg_conv_opts)."""
    import ctypes
    protos = _lib.parse_header()
    assert "dg_conv_set_splitk_workspace" not in protos and "dg_conv_set_tiling" not in protos
    text = _lib.HEADER.read_text()
    m = re.search(r"typedef struct dg_conv_opts \{(.*?)\} dg_conv_opts;", text, flags=re.S)
    fields = [f.strip().split()[-1] for f in m.group(1).split(";") if f.strip()]
    assert fields == [n for n, _ in _lib.ConvOpts._fields_]
    assert ctypes.sizeof(_lib.ConvOpts) == 56
    bad = _lib.ConvOpts(None, 0, 77, -1, -1, 0)
    assert _lib.lib().dg_conv_opts_check(ctypes.byref(bad)) != 0
    assert _lib.lib().dg_conv_opts_check(None) == 0


def test_no_torch_or_cudnn_in_the_abi(built):
    out = subprocess.run(["nm", "-D", "--undefined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for banned in ("cudnn", "cublas", "at::", "c10", "torch"):
        assert banned not in out, banned
    sig = _lib.HEADER.read_text()
    assert "torch" not in sig.lower().replace("pytorch", "")


def test_sass_is_blackwell_native(built):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")     # no legacy mma.sync path


def test_ops_refuse_cpu_tensors():
    import torch
    from discogan_modernized_b200 import ops
    with pytest.raises(ValueError):
        ops._ptr(torch.zeros(4), torch.float32, "x")
    with pytest.raises(ValueError):
        ops._ptr(torch.zeros(4, dtype=torch.float64), None, "x").__class__  # cpu tensor


def test_sass_has_pair_cluster_and_bulk_copy_paths(built):
    """cta_group::2 MMAs / TMA, cluster multicast and the bulk-copy BatchNorm kernels are really in the binary."""
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.5D.2CTA", "UTMALDG.5D.MULTICAST", "UTCBAR.2CTA.MULTICAST", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_launch_plan_query_without_gpu(built):
    """dg_conv_stats_rows is a pure host-side plan query (grid size = rows of the fused-statistics workspace): the
    tile-selection rules can be checked here.  148 SMs are assumed when no device is present."""
    import ctypes
    rows = lambda *a: built.dg_conv_stats_rows(*a, None)
    # deep 16->8 layer of the 512^2 step (B=32): 16 M tiles x 8 N tiles of 256 -> 128 CTAs (64 pairs), not 148 CTAs of
    # 128-wide tiles
    assert rows(0, 32, 8, 8, 2048, 1024) == 128
    # wide layer with plenty of tiles: persistent grid on every SM, launched as 74 CTA pairs
    assert rows(0, 32, 32, 32, 512, 256) == 148
    # <= 128 output channels and many pixel tiles: role-swapped kernel, one CTA per SM
    assert rows(0, 32, 128, 128, 128, 64) == 148
    assert rows(1, 32, 128, 128, 128, 64) == 148
    # tiny problem: one CTA per tile
    assert 0 < rows(0, 2, 4, 4, 128, 64) <= 8
    # the deepest 512^2 layer at B=32 is SM-starved (4 M tiles x 8 N tiles): with a split-K workspace on offer the plan
    # splits (no fused statistics: 0 rows); without one it does not
    sk = _lib.ConvOpts(None, 64 << 20, 0, -1, -1, 0)
    assert built.dg_conv_stats_rows(0, 32, 4, 4, 2048, 2048, ctypes.byref(sk)) == 0
    assert rows(0, 32, 4, 4, 2048, 2048) > 0
