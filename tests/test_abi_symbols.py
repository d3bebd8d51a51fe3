"""The C-ABI library loads and exports every symbol include/vibo_b200.h
declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import vibo_b200
    return vibo_b200


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vibo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vibo_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built._lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vibo_b200.h but not exported"
    assert sorted(built._lib.SIGNATURES) == syms, "ctypes SIGNATURES out of sync with the header"


def test_version_and_error_string(built):
    lib = built._lib.load()
    assert lib.vibo_version() == 100
    # argument validation happens before any CUDA call
    d = built._lib.Desc(10, 5, 99, 2, 0, 0, 0, 0)
    assert lib.vibo_workspace_bytes(ctypes.byref(d)) == 0
    rc = lib.vibo_decode(ctypes.byref(d), None, None, None, None)
    assert rc == -2 and b"ability_dim" in lib.vibo_last_error()
    d = built._lib.Desc(10, 5, 1, 7, 0, 0, 0, 0)
    assert lib.vibo_decode(ctypes.byref(d), None, None, None, None) == -1


def test_no_cpu_fallback(built):
    """CPU tensors are refused, not silently computed elsewhere."""
    import torch
    with pytest.raises(built._lib.ViboError):
        built.kernels.decode(torch.zeros(2, 1), torch.zeros(3, 2), irt_model=2)
    m = built.VIBO_2PL(1, 3, ability_merge="product")
    with pytest.raises(built._lib.ViboError):
        m.fused_elbo(torch.zeros(2, 3, 1), torch.ones(2, 3, 1, dtype=torch.bool))
