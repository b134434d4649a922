#!/usr/bin/env python
"""Throw-away K1 throughput probe (device-resident inputs, native Philox mode) used while tuning.
Not the contract benchmark (that is bench.py)."""
import argparse
import sys, os, time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import capi
from aps_b200.batch import make_batch, make_params


def exp_gradient(L, N, frac_plus, decay):
    xs = np.arange(L) / float(L)
    plus = np.exp(-xs / decay); minus = 0.05 * np.ones_like(xs)
    return N * frac_plus * plus / plus.sum(), N * (1 - frac_plus) * minus / minus.sum()


def host_init_poisson_k1(rng, R, rp, rm):
    L = len(rp)
    cp = rng.poisson(rp, (R, L)); cm = rng.poisson(rm, (R, L))
    tot = cp + cm
    keep_plus = rng.random((R, L)) * tot < cp
    occ = tot > 0
    sig = np.where(keep_plus, 1, -1).astype(np.int8)
    ns = occ.sum(1).astype(np.int32)
    nmax = int(ns.max())
    pos0 = np.zeros((R, nmax), np.int32); sg0 = np.ones((R, nmax), np.int8)
    for r in range(R):
        idx = np.nonzero(occ[r])[0]
        pos0[r, :len(idx)] = idx; sg0[r, :len(idx)] = sig[r, idx]
    return ns, pos0, sg0, nmax


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=4096)
    ap.add_argument("--T", type=float, default=20.0)
    ap.add_argument("--obs_dt", type=float, default=0.1)
    ap.add_argument("--sigma", type=float, default=0.005)
    ap.add_argument("--D", type=float, default=0.02)
    ap.add_argument("--lam", type=float, default=5.0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--record", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max_events", type=int, default=0)
    ap.add_argument("--L", type=int, default=1000)        # other lattice sizes take the runtime-layout kernels (no capacity class)
    ap.add_argument("--N", type=int, default=500)
    a = ap.parse_args()
    lib = capi.load()
    lib.aps_debug_set_k1_threads(a.threads)
    L, R = a.L, a.replicas
    rp, rm = exp_gradient(L, a.N, 0.75, 0.35)
    rm = exp_gradient(L, a.N, 0.75, 0.2)[1]
    rng = np.random.default_rng(0)
    ns, pos0, sg0, nmax = host_init_poisson_k1(rng, R, rp, rm)
    from scipy.ndimage import _filters
    sd = a.sigma / (1.0 / L)
    radius = int(4 * sd + 0.5) if a.sigma > 0 else -1
    w = _filters._gaussian_kernel1d(sd, 0, radius)[::-1].copy() if a.sigma > 0 else np.zeros(1)
    times = np.arange(0.0, a.T, a.obs_dt); M = len(times)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    betas = np.repeat(np.linspace(0, 3, 64), R // 64 + 1)[:R]
    d = dict(times_obs=t(times), weights=t(w), beta=t(betas), n=t(ns), pos0=t(pos0), sigma0=t(sg0),
             seeds=t(np.arange(R, dtype=np.int64) + 1000),
             obs_cp=torch.zeros((R, M, L), dtype=torch.int8, device=dev) if a.record & 1 else None,
             obs_cm=torch.zeros((R, M, L), dtype=torch.int8, device=dev) if a.record & 1 else None,
             obs_pos=torch.zeros((R, M, nmax), dtype=torch.int32, device=dev) if a.record & 2 else None,
             obs_sigma_sum=torch.zeros((R, M), dtype=torch.int32, device=dev),
             obs_m_local=torch.zeros((R, M, L), dtype=torch.float64, device=dev) if a.record & 4 else None,
             n_obs=torch.zeros(R, dtype=torch.int32, device=dev), n_events=torch.zeros(R, dtype=torch.int64, device=dev),
             t_end=torch.zeros(R, dtype=torch.float64, device=dev), status=torch.zeros(R, dtype=torch.int32, device=dev),
             n_guard=torch.zeros(R, dtype=torch.int64, device=dev))
    p = make_params(L, 1, radius, a.D, a.lam, a.T)
    w_host = np.ascontiguousarray(w, dtype=np.float64)
    b, keep = make_batch(R, nmax, M, record=a.record, max_events=a.max_events, weights_host=w_host, **d)
    print(f"R={R} nmax={nmax} mean n={ns.mean():.1f} radius={radius} M={M} smem/replica={lib.aps_replica_smem_bytes(p, nmax)}")
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(a.reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.check(lib.aps_run_philox_device(p, b, st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ev = int(d["n_events"].sum().item())
        print(f"run {rep}: {ms:.2f} ms  events={ev}  {ev / ms * 1e-3:.2f} M events/s  status={torch.bincount(d['status']).tolist()} guard={int(d['n_guard'].sum())}")


if __name__ == "__main__":
    main()
