#!/usr/bin/env python
"""K2 pass timing on one GPU: persistent multi-pass launches vs one launch per pass, at the configuration BASELINE
config 5 names (L = 2^26, local field sigma = 5 sites, dt = 0.005; global field, dt = 0.02) and at L = 2^30.
Prints one JSON line per case: us per pass, GB/s (2 B per site-visit), fraction of the measured HBM copy peak,
particle attempts per second, simulated time units per wall second."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200.sublattice import SublatticeLattice

ap = argparse.ArgumentParser()
ap.add_argument("--logL", type=int, nargs="+", default=[26, 30])
ap.add_argument("--passes", type=int, default=200)
ap.add_argument("--case", default="", help="substring filter on the case names")
ap.add_argument("--legacy", action="store_true", help="also time the persistent multi-pass launch (grid barrier per pass)")
ap.add_argument("--ctas", type=int, default=0, help="A/B: persistent CTAs per SM (aps_debug_set_k2_ctas_per_sm); ring depth via APS_K2_STAGES")
a = ap.parse_args()
if a.ctas:
    from aps_b200 import capi
    capi.load().aps_debug_set_k2_ctas_per_sm(a.ctas)
peak = 6548.2
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
CASES = [("local sigma=5 dt=0.005", 5.0, 0.005), ("local sigma=5 dt=0.0025", 5.0, 0.0025), ("global dt=0.02", None, 0.02),
         ("global dt=0.005", None, 0.005), ("global dt=0.0025", None, 0.0025)]
for logL in a.logL:
    L = 1 << logL
    passes = a.passes if logL <= 28 else max(20, a.passes // 8)
    for name, sigma, dt in CASES:
        if a.case not in name:
            continue
        for persistent in ([False, True] if a.legacy else [False]):
            lat = SublatticeLattice(L, D=0.02, lam=5.0, beta=2.0, dt=dt, sigma_sites=sigma, seed=0, single_rank=True,
                                    persistent=persistent)
            lat.init_random(0.5, 0.5)
            lat.run_passes(6)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lat.run_passes(passes); e1.record(); torch.cuda.synchronize()
            lat.check()
            us = 1e3 * e0.elapsed_time(e1) / passes
            gbs = 2.0 * L / (us * 1e-6) / 1e9
            print(json.dumps(dict(logL=logL, case=name, launch="persistent" if persistent else "per-pass", us_per_pass=round(us, 2),
                                  GBs=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak, 3), attempts_per_s=lat.n_particles / (us * 1e-6),
                                  sim_time_per_s=0.5 * dt / (us * 1e-6), trials_per_half=round(lat.rates.mu, 2))), flush=True)
            del lat
            torch.cuda.empty_cache()
