"""CPU-side checks of the drop-in boundary: the library builds, loads, exports every symbol that
include/aps.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from aps_b200 import capi
from aps_b200.batch import make_batch, make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aps.h")).read() + open(os.path.join(ROOT, "include", "aps_pde.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aps_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from aps_b200 import build
    build.build()
    return capi.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aps.h but not exported"
        assert n in capi.SYMBOLS, f"{n} has no ctypes prototype in capi.py"
    assert sorted(capi.SYMBOLS) == names
    assert lib.aps_abi_version() == capi.ABI_VERSION == 5


def test_struct_layout_matches_header(lib):
    # sizes the C compiler gives the two descriptors (x86-64 SysV): 4 int32 + 6 double; 4 int32 + 3 int64 + 36 pointers + exit_cap
    assert C.sizeof(capi.ApsParams) == 64
    assert C.sizeof(capi.ApsBatch) == 16 + 24 + 36 * 8 + 8 + 24           # + flip_tab, flip_G, weights_host (ABI 5)
    assert C.sizeof(capi.ApsPdeArgs) == 8 * 4 + 8 + 3 * 8 + 18 * 8      # include/aps_pde.h
    assert C.sizeof(capi.ApsProfileArgs) == 6 * 4 + 8 + 8 * 8 + 8 and C.sizeof(capi.ApsHistArgs) == 8 * 4 + 2 * 8 + 7 * 8
    assert C.sizeof(capi.ApsK2Multi) == 4 * 4 + 3 * 8 + 3 * 8 + 8 * 8 and lib.aps_k2_peer_region_bytes() == 320 + 4 * 65536


def test_invalid_arguments_are_rejected(lib):
    p = make_params(0, 1, -1, 0.0, 1.0, 1.0)
    b, _ = make_batch(1, 4, 1)
    assert lib.aps_run_replay_host(p, b) == capi.APS_ERR_INVALID
    assert b"L" in lib.aps_last_error()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.aps_device_count() == 0
    t = np.zeros(1); beta = np.ones(1); n = np.array([1], np.int32)
    pos = np.zeros((1, 1), np.int32); sg = np.ones((1, 1), np.int8); seeds = np.zeros(1, np.uint64)
    p = make_params(8, 1, -1, 0.0, 1.0, 1.0)
    b, keep = make_batch(1, 1, 1, times_obs=t, beta=beta, n=n, pos0=pos, sigma0=sg, seeds=seeds)
    assert lib.aps_run_philox_host(p, b) == capi.APS_ERR_NO_DEVICE
    with pytest.raises(capi.ApsError):
        capi.check(lib.aps_run_philox_host(p, b), "run")
