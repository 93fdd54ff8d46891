"""CPU: the C-ABI library loads and exports every symbol include/rlg_b200.h declares; host-side argument
validation returns the documented negative codes without touching a GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rlg_b200.h")


@pytest.fixture(scope="module")
def lib():
    import importlib
    build = importlib.import_module("gan-rl_3d_b200.build")
    build.build_library()
    _lib = importlib.import_module("gan-rl_3d_b200._lib")
    return _lib.load(), _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rlg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    cdll, _lib = lib
    names = declared_functions()
    assert {"rlg_chamfer_fwd", "rlg_chamfer_bwd", "rlg_encoder_fwd", "rlg_chamfer_ws_bytes"} <= set(names)
    for name in names:
        assert hasattr(cdll, name), f"{name} declared in include/rlg_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), "ctypes table and header disagree"


def test_version_and_ws_bytes(lib):
    cdll, _ = lib
    assert cdll.rlg_version() == 7
    # per point: 8-byte key + 4-byte runner-up value + (tensor sweep) 4-byte runner-up group + 4-byte third value; + counters
    big = cdll.rlg_chamfer_ws_bytes(32, 2048, 2048)
    assert 20 * 32 * 4096 < big <= 20 * 32 * 4096 + 16384 and big % 256 == 0
    assert 0 < cdll.rlg_chamfer_ws_bytes(1, 1, 1) <= 4096
    assert cdll.rlg_chamfer_ws_bytes(1, 0, 5) == 0


def test_argument_errors_are_negative_codes_with_messages(lib):
    cdll, _lib = lib
    # empty clouds: the reference raises IndexError; here RLG_ERR_BAD_SHAPE before any launch
    rc = cdll.rlg_chamfer_fwd(None, None, 2, 0, 5, None, None, None, None, None, None, None, 0, 0, None)
    assert rc == -2 and b"bad shape" in cdll.rlg_last_error()
    rc = cdll.rlg_chamfer_fwd(None, None, 2, 4, 5, None, None, None, None, None, None, None, 0, 0, None)
    assert rc == -1
    rc = cdll.rlg_chamfer_bwd(*([None] * 8), 1, 3, 0, None, None, 0, None)
    assert rc == -2
    with pytest.raises(_lib.RlgError) as ei:
        _lib.check("rlg_chamfer_bwd", rc)
    assert ei.value.code == -2
    # the pair index is a grid dimension (like the forward's): at most 65535 pairs per call
    assert cdll.rlg_chamfer_bwd(*([None] * 8), 65536, 3, 3, None, None, 0, None) == -5
    # unknown backward flag bits are rejected; the reproducible variant insists on its workspace (24 bytes per point)
    assert cdll.rlg_chamfer_bwd(*([None] * 8), 1, 3, 3, None, None, 6, None) == -4
    assert cdll.rlg_chamfer_bwd_ws_bytes(2, 5, 7) == 24 * 2 * 12 and cdll.rlg_chamfer_bwd_ws_bytes(0, 5, 7) == 0
    one = ctypes.c_void_p(256)      # non-null, never dereferenced: argument checks come first
    assert cdll.rlg_chamfer_bwd_det(*([one] * 8), 1, 3, 3, one, one, None, 0, 0, None) == -1
    assert cdll.rlg_chamfer_bwd_det(*([one] * 8), 1, 3, 3, one, one, one, 8, 0, None) == -3
    # B == 0 is a no-op, like an empty batch through the reference
    assert cdll.rlg_chamfer_fwd(None, None, 0, 4, 5, None, None, None, None, None, None, None, 0, 0, None) == 0
    layer = (_lib.RlgLayer * 1)()
    layer[0].c_in, layer[0].c_out = 4, 8
    assert cdll.rlg_encoder_fwd(1, 1, 16, layer, 1, 1, None, 1, 256, None) == -4   # layer 0 must be 3 -> c


def test_missing_library_is_loud(monkeypatch):
    import importlib
    _lib = importlib.import_module("gan-rl_3d_b200._lib")
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librlg_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()
