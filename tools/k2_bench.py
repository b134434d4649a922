#!/usr/bin/env python
"""K2 throughput / HBM-roofline probe (device-resident lattice)."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200.sublattice import SublatticeLattice

ap = argparse.ArgumentParser()
ap.add_argument("--logL", type=int, default=30)
ap.add_argument("--beta", type=float, default=2.0)
ap.add_argument("--dt", type=float, default=0.02)
ap.add_argument("--sigma", type=float, default=5.0)
ap.add_argument("--passes", type=int, default=40)
ap.add_argument("--D", type=float, default=0.02)
ap.add_argument("--lam", type=float, default=5.0)
ap.add_argument("--ctas-per-sm", type=int, default=0, help="persistent CTAs per SM (debug hook; 0 = library default)")
ap.add_argument("--stash-cap", type=int, default=0, help="local-field stash capacity (debug hook; 0 = automatic)")
a = ap.parse_args()
if a.ctas_per_sm or a.stash_cap:
    from aps_b200 import capi
    if a.ctas_per_sm: capi.load().aps_debug_set_k2_ctas_per_sm(a.ctas_per_sm)
    if a.stash_cap: capi.load().aps_debug_set_k2_stash_cap(a.stash_cap)
L = 1 << a.logL
lat = SublatticeLattice(L, D=a.D, lam=a.lam, beta=a.beta, dt=a.dt, sigma_sites=a.sigma if a.sigma > 0 else None, seed=0)
lat.init_random(0.5, 0.5)
lat.run_passes(6)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lat.run_passes(a.passes); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.passes
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
gbs = 2.0 * L / (ms * 1e-3) / 1e9
print(json.dumps(dict(ctas_per_sm=a.ctas_per_sm, stash_cap=a.stash_cap, L=L, beta=a.beta, dt=a.dt, sigma=a.sigma, mu=lat.rates.mu, ms_per_pass=ms, site_visits_per_s=L / (ms * 1e-3),
                      particle_attempts_per_s=lat.n_particles / (ms * 1e-3), hbm_gbs=gbs, frac_of_measured_peak=gbs / peak,
                      trials_per_s=lat.rates.mu * (L / 64) / (ms * 1e-3))))
