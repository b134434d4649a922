"""CPU stand-in for sublattice.CudaK2Backend (TEST HELPER): torch CPU tensors + the oracle's K2."""
import numpy as np
import torch

from aps_b200.capi import ApsK2Rates
from oracle import oracle


class OracleK2Backend:
    name = "oracle"
    dev = torch.device("cpu")

    def __init__(self):
        self.lib = oracle.load()

    def zeros_u8(self, n): return torch.zeros(n, dtype=torch.uint8)
    def zeros_i64(self, n): return torch.zeros(n, dtype=torch.int64)
    def from_numpy(self, a): return torch.from_numpy(np.ascontiguousarray(a).copy())
    def ptr(self, t): return t.data_ptr()

    def rates(self, D, lam, beta, dt):
        r = ApsK2Rates()
        assert self.lib.aps_oracle_k2_rates(D, lam, beta, dt, r) == 0
        return r

    def flip_table(self, rates):
        import ctypes
        t = np.zeros(2 * 1025, np.uint32)
        self.lib.aps_oracle_k2_flip_table(ctypes.addressof(rates), t.ctypes.data)
        return torch.from_numpy(t.view(np.int32).copy())

    def run(self, args, n_passes):
        assert self.lib.aps_oracle_k2_run(args, n_passes) == 0

    def init(self, state, L, off, seed, density, frac_plus):
        self.lib.aps_oracle_k2_init(state.data_ptr(), L, off, seed, density, frac_plus)

    def profile(self, state, L, off, L_global, nbins):
        s = state.numpy()
        g = off + np.arange(L, dtype=np.int64)
        b = (g * nbins) // L_global
        cp = np.bincount(b[s == 1], minlength=nbins).astype(np.int64)
        cm = np.bincount(b[s == 2], minlength=nbins).astype(np.int64)
        return torch.from_numpy(cp), torch.from_numpy(cm)

    def count(self, view):
        return int((view == 1).sum()), int((view == 2).sum())
