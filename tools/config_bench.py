#!/usr/bin/env python
"""End-to-end timings of BASELINE configs 3 and 4 through the launcher (host parameters -> host results), for the
measurement table in DESIGN.md.  bench.py stays the contract benchmark (config 2)."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import launcher as la

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--scale", type=float, default=1.0, help="fraction of the replicas per grid point")
a = ap.parse_args()
base = dict(L=1000, xlim=1, flip_rate_fn=None, scale_rates=False, minus_anchor=True, periodic=False, anchor_positions=None,
            site_capacity=1, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)

def timed(fn):
    fn(); torch.cuda.synchronize()                         # warm-up (context, cuFFT plans, allocator)
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    return out, time.perf_counter() - t0

if a.config == 3:       # local_structure.py:675-726: N = 900 fixed, D = 0.05, lambda = 5, sigma = 0.005, T = 40, obs_dt = 1, 64 beta x 64 runs
    ps = dict(base, rate_diffusion=0.05, rate_active=5, init="fixed", N=900, local_kernel_sigma=0.005)
    runs = max(1, int(64 * a.scale))
    betas = list(np.linspace(0, 3, 64))
    out, dt = timed(lambda: la.sweep_betas_for_structures(betas, runs, ps, {}, dict(T=40.0, obs_dt=1.0), keep_raw=False))
    ev = int(sum(int(out[b]["n_events"].sum()) for b in betas))
    print(json.dumps(dict(config=3, replicas=64 * runs, seconds=dt, events=ev, events_per_s=ev / dt,
                          dominant_k=[out[b]["dominant_k_mode"] for b in betas[::16]])))
else:                    # double_sweep.py:666-715: 16 beta x 16 densities x 128 runs, D = 0.005, lambda = 10, sigma = 0.02 (r = 80), T = 10
    ps = dict(base, rate_diffusion=0.005, rate_active=10, local_kernel_sigma=0.02)
    runs = max(1, int(128 * a.scale))
    Ns = [int(v) for v in np.linspace(50, 950, 16)]
    betas = list(np.linspace(0, 3, 16))
    spec = la.build_double_sweep_spec(Ns, betas, runs, ps, dict(T=10.0, obs_dt=0.1), frac_plus=0.75, decay_plus=0.2)
    def go():
        return la.run_ensemble(spec, want_profiles=False)
    res, dt = timed(go)
    ev = int(res.n_events.sum())
    print(json.dumps(dict(config=4, replicas=len(spec.betas), seconds=dt, events=ev, events_per_s=ev / dt,
                          status_ok=bool((res.status == 0).all()))))
