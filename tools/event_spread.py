#!/usr/bin/env python
"""Spread of the per-replica event counts of config 2 (what bounds a single-wave K1 launch: its slowest replica)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as B
from aps_b200 import launcher as la
betas = np.linspace(0, 3, B.N_BETA)
out = la.sweep_over_betas(betas, B.REPS_PER_BETA, B.PS_KWARGS, B.init_kwargs(), dict(B.RUN_KWARGS, T=20.0), base_seed=1, want_profiles=False)
ev = out["n_events"].astype(float)
print("mean", ev.mean(), "max", ev.max(), "min", ev.min(), "mean/max", ev.mean() / ev.max())
print("per-beta mean (every 8th):", np.round(ev.mean(axis=1)[::8]).tolist())
print("per-beta max  (every 8th):", ev.max(axis=1)[::8].tolist())
q = np.quantile(ev, [0.5, 0.9, 0.99, 0.999])
print("quantiles 50/90/99/99.9 %:", q.tolist())
