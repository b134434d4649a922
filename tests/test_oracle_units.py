"""Unit-level pins of the oracle's building blocks against numpy / scipy / published vectors."""
import ctypes as C

import numpy as np
import pytest

from aps_b200.batch import make_params
from oracle import oracle


def ulp_diff(a, b):
    ia = np.asarray(a, np.float64).view(np.int64)
    ib = np.asarray(b, np.float64).view(np.int64)
    return np.abs(ia - ib)


def test_exp_within_one_ulp_of_libm():
    lib = oracle.load()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-12, 12, 200000), rng.uniform(-1e-3, 1e-3, 20000),
                         rng.uniform(-700, 700, 20000), [0.0, 1.0, -1.0, 0.5 * np.log(2), 1e-30, -1e-30]])
    got = np.array([lib.aps_oracle_exp(float(x)) for x in xs])
    assert ulp_diff(got, np.exp(xs)).max() <= 1
    assert lib.aps_oracle_exp(0.0) == 1.0
    assert lib.aps_oracle_exp(1000.0) == np.inf and lib.aps_oracle_exp(-1000.0) == 0.0


def test_log_within_one_ulp_of_libm():
    lib = oracle.load()
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(0, 1, 200000), 1 - rng.uniform(0, 1e-6, 20000),
                         np.exp(rng.uniform(-50, 50, 20000)), [1.0, 0.5, 2.0, 2.0 ** -53]])
    xs = xs[xs > 0]
    got = np.array([lib.aps_oracle_log(float(x)) for x in xs])
    assert ulp_diff(got, np.log(xs)).max() <= 1
    assert lib.aps_oracle_log(1.0) == 0.0


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    lib = oracle.load()
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kats:
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); o = np.zeros(4, np.uint32)
        lib.aps_oracle_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert o.tolist() == want


def test_native_total_follows_its_written_definition():
    """Native-mode total rate (include/aps_math.h): chunks of CS = 16 * 2^k rates (k smallest with <= 32 chunks, rates beyond n are
    0.0), chunk sum = T16 of the first 16 rates plus, left to right, T16 of every further block of 16 starting below n, T16 = adjacent
    pairwise tree; R = last lane of a 32-lane Hillis-Steele scan of the chunk sums.  Restated here with numpy float64 scalars and
    compared bit for bit; the kernels are compared with the oracle in the GPU tests."""
    import ctypes
    lib = oracle.load()
    lib.aps_oracle_native_total.restype = ctypes.c_double
    lib.aps_oracle_native_total.argtypes = [ctypes.c_void_p, ctypes.c_int]

    def t16(r, lo, n):
        a = [np.float64(r[lo + k]) if lo + k < n else np.float64(0.0) for k in range(16)]
        w = 1
        while w < 16:
            for k in range(0, 16, 2 * w):
                a[k] = a[k] + a[k + w]
            w *= 2
        return a[0]

    def total(r):
        n = len(r)
        sh = 4
        while ((n + (1 << sh) - 1) >> sh) > 32:
            sh += 1
        cs = 1 << sh
        v = []
        for j in range(32):
            lo = j << sh
            if lo >= n:
                v.append(np.float64(0.0)); continue
            c = t16(r, lo, n)
            b = lo + 16
            while b < lo + cs and b < n:
                c = c + t16(r, b, n); b += 16
            v.append(c)
        o = 1
        while o < 32:
            v = [v[l] + v[l - o] if l >= o else v[l] for l in range(32)]
            o *= 2
        return v[31]

    g = np.random.default_rng(5)
    for n in [1, 2, 15, 16, 17, 31, 33, 369, 488, 511, 512, 513, 700, 968, 1024, 1500, 2048, 3000]:
        r = np.ascontiguousarray(g.random(n) * 7.0 + 0.01)
        got = lib.aps_oracle_native_total(r.ctypes.data, n)
        want = total(r)
        assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64), n
        assert abs(got - r.sum()) <= 1e-12 * r.sum()


def test_pairwise_sum_equals_numpy_sum():
    """R = rates.sum() (CLASS.py:352) is numpy's pairwise summation."""
    lib = oracle.load()
    rng = np.random.default_rng(3)
    for n in list(range(1, 300)) + [343, 377, 500, 750, 900, 1000, 1024, 4097]:
        a = rng.uniform(0.1, 7.0, n)
        got = lib.aps_oracle_pairwise_sum(a.ctypes.data, n)
        assert got == a.sum(), n


@pytest.mark.parametrize("L,sigma_grid", [(1000, 5.0), (1000, 2.0), (1000, 20.0), (200, 300.0), (1000, 0.1),
                                          (1000, 0.5), (100, 30.0), (16, 1.6), (7, 3.3)])
def test_m_field_equals_scipy_gaussian_filter(L, sigma_grid):
    """compute_local_m_field (CLASS.py:216-246) restated == scipy.ndimage.gaussian_filter1d path."""
    from scipy.ndimage import _filters, gaussian_filter1d

    lib = oracle.load()
    rng = np.random.default_rng(4)
    cp = rng.integers(0, 3, L).astype(np.int32)
    cm = rng.integers(0, 3, L).astype(np.int32)
    cp[rng.random(L) < 0.4] = 0
    cm[rng.random(L) < 0.6] = 0
    s = cp.astype(float) - cm.astype(float)
    tot = cp.astype(float) + cm.astype(float)
    sc = gaussian_filter1d(s, sigma=sigma_grid, mode="reflect")
    tc = gaussian_filter1d(tot, sigma=sigma_grid, mode="reflect")
    want = np.zeros(L)
    mask = tc > 0
    want[mask] = sc[mask] / tc[mask]
    want = np.clip(want, -1, 1)
    r = int(4.0 * sigma_grid + 0.5)
    w = _filters._gaussian_kernel1d(sigma_grid, 0, r)[::-1].copy()
    out = np.zeros(L)
    p = make_params(L, 3, r, 0, 0, 1)
    assert lib.aps_oracle_m_field(p, w.ctypes.data, cp.ctypes.data, cm.ctypes.data, out.ctypes.data) == 0
    assert np.array_equal(out, want)


def test_m_field_global():
    lib = oracle.load()
    cp = np.array([1, 0, 2, 0, 1], np.int32); cm = np.array([0, 1, 0, 0, 2], np.int32)
    out = np.zeros(5)
    p = make_params(5, 3, -1, 0, 0, 1)
    lib.aps_oracle_m_field(p, None, cp.ctypes.data, cm.ctypes.data, out.ctypes.data)
    assert np.array_equal(out, np.full(5, (4 - 3) / 7.0))
