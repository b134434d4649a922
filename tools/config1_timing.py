#!/usr/bin/env python
"""BASELINE config 1: PARTICLE_solver_BIOLOGY_EXCLUSION.py single run (script defaults, :55-97) through the drop-in class.
Wall time of ps.run(T=20, obs_dt=0.5, record_fft=True, record_var=True) in replay mode (seeded numpy Generator: same
trajectory as the reference) and in native mode.  Reference: ~3.6e4 events at ~298 us/event = ~10.7 s."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
from PARTICLE_solver_CLASS import ParticleSystem, PhiloxRNG

kw = dict(L=1000, xlim=1, rate_diffusion=0, rate_active=5, beta=0.7, flip_rate_fn=None, init="fixed", N=750, scale_rates=False,
          local_kernel_sigma=0.002, minus_anchor=True, periodic=False, immobilize_when_anchored=True, anchor_radius=0.003,
          anchor_positions=None, site_capacity=3, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)
res = {}
for mode, mk in [("replay(default_rng(0))", lambda: np.random.default_rng(0)), ("native(PhiloxRNG(0))", lambda: PhiloxRNG(0))]:
    for rep in range(3):
        ps = ParticleSystem(rng=mk(), **kw)
        t0 = time.perf_counter()
        out = ps.run(T=20, obs_dt=0.5, record_fft=True, record_var=True)
        dt = time.perf_counter() - t0
    res[mode] = dict(seconds=dt, **ps.last_run_info)
# sweep_beta single run (K=1, D>0: speculative chunks with rewinds)
from aps_b200.launcher import make_exp_gradient
kw2 = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, beta=1.5, flip_rate_fn=None, init="poisson", N=500, scale_rates=False,
           local_kernel_sigma=0.005, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False,
           k_on=0, k_off=0, k_exit=0,
           rho0_plus=make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)[0],
           rho0_minus=make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.2, anchor_positions=None)[1])
for mode, mk in [("sweep_beta run replay", lambda: np.random.default_rng(1)), ("sweep_beta run native", lambda: PhiloxRNG(1))]:
    for rep in range(2):
        ps = ParticleSystem(rng=mk(), **kw2)
        t0 = time.perf_counter()
        out = ps.run(T=20, obs_dt=0.1, record_fft=True, record_var=True)
        dt = time.perf_counter() - t0
    res[mode] = dict(seconds=dt, **ps.last_run_info)
print(json.dumps(res, indent=1))
