"""The C-ABI library builds, loads and exports every entry point include/audiopure_b200.h declares.
No GPU and no compute calls here."""

import os
import re
import subprocess

import pytest

from audiopure_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "audiopure_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ap_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_for_sm_100a():
    path = build.build()
    assert os.path.exists(path)
    sass = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_every_declared_entry_point_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 18
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.ap_abi_version() == _lib.AP_ABI_VERSION
    # argument validation happens before any CUDA call, so it works without a device
    assert lib.ap_workspace_bytes(None, 1, 16000) == 0
    rc = lib.ap_vote_counts(None, 1, 10, None, None)
    assert rc != 0 and b"null" in lib.ap_last_error()
    with pytest.raises(_lib.AudioPureError):
        _lib.check(lib.ap_debug_gemm(None, None, None, 64, None))


def test_kernels_use_blackwell_tensor_and_tma_instructions():
    """SASS evidence: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG."""
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")  # no legacy mma.sync path


def test_no_cpu_fallback():
    import torch

    import audiopure_b200 as ap

    m = ap.WaveNet_Speech_Commands(res_channels=256, skip_channels=256, num_res_layers=2, dilation_cycle=2)
    with pytest.raises(_lib.AudioPureError):
        m((torch.zeros(1, 1, 256), torch.zeros(1, 1)))
    with pytest.raises(_lib.AudioPureError):
        ap.LogMelSpectrogram()(torch.zeros(1, 1, 16000))
