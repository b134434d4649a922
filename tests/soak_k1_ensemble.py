#!/usr/bin/env python
"""Manual soak (not collected by pytest): the WHOLE config-2 ensemble (4096 replicas, native Philox mode, T given on the
command line) on the GPU against the CPU oracle running the same Philox streams, every replica, bitwise.
Usage on a GPU box:  python tests/soak_k1_ensemble.py 5.0     (about a minute of oracle time on 16 cores)."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np, torch
from aps_b200 import launcher as la
from aps_b200.batch import make_params
from common import HostRun, run_oracle

T = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
PS = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, flip_rate_fn=None, init="poisson", N=500, scale_rates=False,
          local_kernel_sigma=0.005, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False)
g = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
g2 = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.2, anchor_positions=None)
spec = la.build_beta_sweep_spec(np.linspace(0, 3, 64), 64, PS, dict(rho0_plus=g[0], rho0_minus=g2[1]), dict(T=T, obs_dt=0.1), base_seed=21)
ens = la.DeviceEnsemble(spec, 0, 4096)
ens.init_particles()
rb = ens.rb
rb.run_philox(); torch.cuda.synchronize()
n = ens.n.cpu().numpy(); pos0 = ens.pos0.cpu().numpy(); sg0 = ens.sigma0.cpu().numpy()
M = rb.M
hr = HostRun(1000, ens.n_max, M, n, pos0, sg0, np.asarray(spec.betas), ens.times_obs, ens.mp["weights"],
             seeds=np.asarray(spec.seeds, dtype=np.uint64), record=3)
params = make_params(1000, 1, ens.mp["radius"], ens.mp["D"], ens.mp["lam"], T, 0)
t0 = time.perf_counter()
run_oracle(params, hr, mode=1, threads=os.cpu_count() or 4)
print(f"oracle: {hr.n_events.sum()} events in {time.perf_counter() - t0:.1f} s")
bad = 0
for name, dev in [("n_events", rb.n_events), ("n_obs", rb.n_obs), ("status", rb.status), ("pos_end", rb.pos_end), ("sigma_end", rb.sigma_end),
                  ("obs_cp", rb.obs_cp), ("obs_cm", rb.obs_cm), ("obs_pos", rb.obs_pos), ("obs_sigma_sum", rb.obs_sigma_sum)]:
    a, b = dev.cpu().numpy().reshape(getattr(hr, name).shape), getattr(hr, name)
    if name in ("pos_end", "sigma_end", "obs_pos"):                 # ragged rows: compare the first n entries of every replica
        live = np.arange(ens.n_max)[None, :] < n[:, None]
        live = live if a.ndim == 2 else np.broadcast_to(live[:, None, :], a.shape)
        a, b = np.where(live, a, 0), np.where(live, b, 0)
    ok = np.array_equal(a, b)
    bad += not ok
    print(f"{name:14s} {'identical' if ok else 'DIFFERENT'}")
ok = np.array_equal(rb.t_end.cpu().numpy().view(np.uint64), hr.t_end.view(np.uint64))
bad += not ok
print(f"{'t_end (bits)':14s} {'identical' if ok else 'DIFFERENT'};  guard-band slow path taken {int(rb.n_guard.sum())} times in {int(rb.n_events.sum())} events")
sys.exit(1 if bad else 0)
