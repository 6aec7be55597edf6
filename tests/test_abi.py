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
    assert len(protos["dg_conv4x4s2_wgrad"][1]) == 12


def test_library_exports_all_symbols(built):
    for name in _lib.parse_header():
        assert hasattr(built, name), name
    assert built.dg_version() >= 1
    assert built.dg_launch_count() == 0


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
