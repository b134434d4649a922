#!/usr/bin/env python
"""Timing of the batched IMEX PDE stepper on the reference's sweep shapes (IMEX_PDE_solver_run_sweep.py: 33 runs of
80 000 steps at L = 1000, 1000 tracers each; the reference needs ~1.5 ms per step and run on one CPU core)."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200.imex_pde import IMEXPDE, solve_many

ap = argparse.ArgumentParser()
ap.add_argument("--runs", type=int, default=33)
ap.add_argument("--T", type=float, default=40.0)
ap.add_argument("--L", type=int, default=1000)
ap.add_argument("--sigma", type=float, default=0.005)
ap.add_argument("--gamma", type=float, default=0.2)
ap.add_argument("--tracers", type=int, default=1000)
a = ap.parse_args()
betas = np.linspace(0, 3, a.runs)
solvers = []
for i, b in enumerate(betas):
    s = IMEXPDE(L=a.L, T=a.T, dt=5e-4, gamma=a.gamma, lam=0.6, beta=float(b), bc="periodic", active_model="bidirectional",
                gaussian_kernel=True, kernel_sigma=a.sigma, snapshot_interval=50, outdir="/tmp/imex_bench", seed=i)
    s.initialize(mode="homogeneous", rho0=1.0, noise=0.3, n_tracers=a.tracers)
    solvers.append(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
solve_many(solvers)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = solvers[0].nsteps * a.runs
print(json.dumps(dict(runs=a.runs, L=a.L, nsteps=solvers[0].nsteps, kernel_sigma=a.sigma, kernel_radius=solvers[0].kernel_radius,
                      gamma=a.gamma, tracers=a.tracers, seconds=dt, us_per_step_per_run=dt / solvers[0].nsteps * 1e6,
                      run_steps_per_s=steps / dt, m_end=[float(s.m_series[-1]) for s in solvers[:: max(1, a.runs // 6)]])))
