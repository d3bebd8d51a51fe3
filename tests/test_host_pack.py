"""Host-side packing of the reference's (response, mask) rows into the 1 B/cell transfer format
(vibo_pack_host, csrc/vibo_hostpack.cpp): host-only code of the library, checked on CPU against numpy."""
import ctypes as C

import numpy as np
import pytest

import vibo_b200
from vibo_b200 import _lib, kernels as K


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.mark.parametrize("P,I", [(1, 1), (3, 5), (257, 95), (4099, 1000), (700, 1531)])
def test_pack_host_matches_numpy(lib, P, I):
    rng = np.random.default_rng(P * 31 + I)
    resp = (rng.random((P, I)) < 0.5).astype(np.float32)
    mask = (rng.random((P, I)) >= 0.2).astype(np.uint8)
    resp[mask == 0] = -1.0          # the reference's MISSING_DATA marker (src/config.py:14)
    out = np.full((P, I), 7, dtype=np.int8)
    desc = K.make_desc(P, I, 1, 2, False)
    rc = lib.vibo_pack_host(C.byref(desc), resp.ctypes.data, mask.ctypes.data, out.ctypes.data)
    assert rc == 0
    want = np.where(mask != 0, (resp > 0.5).astype(np.int8), np.int8(-1))
    assert np.array_equal(out, want)


def test_pack_host_many_tasks_and_threads(lib):
    """More cells than one task (2^18): the pool splits the range; every byte is written exactly once."""
    assert lib.vibo_host_threads() >= 1
    n = (1 << 20) + 12345
    rng = np.random.default_rng(0)
    resp = (rng.random(n) < 0.3).astype(np.float32)
    mask = (rng.random(n) >= 0.05).astype(np.uint8)
    out = np.full(n, 9, dtype=np.int8)
    desc = K.make_desc(1, n, 1, 2, False)
    for _ in range(3):   # the pool is persistent: repeated parallel regions
        out[:] = 9
        assert lib.vibo_pack_host(C.byref(desc), resp.ctypes.data, mask.ctypes.data, out.ctypes.data) == 0
        assert np.array_equal(out, np.where(mask != 0, (resp > 0.5).astype(np.int8), np.int8(-1)))


def test_pack_host_bad_arguments(lib):
    desc = K.make_desc(4, 4, 1, 2, False)
    assert lib.vibo_pack_host(C.byref(desc), None, None, None) != 0
    assert b"vibo_pack_host" in lib.vibo_last_error()


def test_pack_rows_host_wrapper():
    import torch
    resp = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, 1.0]]).unsqueeze(2)
    mask = torch.tensor([[1, 1, 0], [1, 0, 1]], dtype=torch.bool).unsqueeze(2)
    out = vibo_b200.kernels.pack_rows_host(resp, mask)
    assert out.dtype == torch.int8 and out.tolist() == [[1, 0, -1], [0, -1, 1]]
