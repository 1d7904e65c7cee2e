"""The C-ABI shared library loads and exports every symbol include/cmtcoop_b200.h declares.
No compute is attempted without a GPU: on this CPU-only container every launch entry must refuse
with CMT_ERR_ARCH (there is no fallback implementation to fall back to)."""
import ctypes
import os

import pytest

from cmtcoop_b200 import _lib


def test_header_and_bindings_agree():
    declared = set(_lib.header_symbols())
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert len(declared) >= 13


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.header_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    lib = _lib.load()
    assert lib.cmt_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_sass_is_blackwell_native():
    """tcgen05 / TMA evidence in the built library (cuobjdump is part of the CUDA toolkit)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_entries_refuse_without_sm100():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    lib = _lib.load()
    assert lib.cmt_check_device(0) == -3
    rc = lib.cmt_coop_max(None, None, None, 0, None)
    assert rc == -3 and "no CUDA device" in _lib.last_error() or "sm_100" in _lib.last_error()
    rc = lib.cmt_gemm_bias_act(None, None, None, None, 1, 1, 8, 8, 8, 1, 1, 0, 1, 0, 0, 0, 1.0, 0, 1, 1, None, None)
    assert rc == -3
