#!/usr/bin/env python
"""Full-length reference pins: run the UNMODIFIED reference (`/root/reference/PARTICLE_solver_CLASS.py`,
matplotlib/vispy stubbed) through its public API at the full length of the BASELINE configurations and store,
per run, the seed plus digests of every returned array -> tests/golden/full_length.json (+ small .npz of rows).

A run of the reference is a pure function of (constructor keywords, run keywords, seed of the numpy Generator passed
through the constructor's `rng=` seam, PARTICLE_solver_CLASS.py:26,75-78), and the drop-in consumes that Generator in
the reference's call order, so seed + digests pin the whole trajectory: >= 10^6 events in total
(SURVEY 7.1 gate).  The only wrapper is a call counter that forwards every call unchanged.

Cases: config 1 (T=20), config 2 (T=20) at 8 betas x 6 seeds, config 3 (T=40) at 2 betas, config 4 (T=10, r=80) at
2 (N, beta) points, the sigma=0 global-field sweep point of sweep_beta_2 (T=20).

Usage: python tools/gen_golden_full.py [--procs 8]      (build container only: needs /root/reference)
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

BASE = dict(flip_rate_fn=None, minus_anchor=True, periodic=False, immobilize_when_anchored=True, anchor_radius=0.003,
            anchor_positions=None, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)


class CountingRNG:
    """Forwards every call to the numpy Generator unchanged; counts events (one exponential per event, CLASS.py:358)."""

    def __init__(self, gen):
        self.gen = gen
        self.n_exp = 0
        self.n_random = 0

    def exponential(self, *a, **k):
        self.n_exp += 1
        return self.gen.exponential(*a, **k)

    def choice(self, *a, **k):
        return self.gen.choice(*a, **k)

    def random(self, *a, **k):
        self.n_random += 1
        return self.gen.random(*a, **k)

    def poisson(self, *a, **k):
        return self.gen.poisson(*a, **k)


def full_cases():
    """name -> dict(ps=..., profile=..., run=..., seed=...) in the format of tools/gen_golden.py cases()."""
    c = {}
    for s in (0, 1):
        c[f"c1_T20_s{s}"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0, rate_active=5, beta=0.7, init="fixed", N=750,
                                         scale_rates=False, local_kernel_sigma=0.002, site_capacity=3),
                                 run=dict(T=20.0, obs_dt=0.5, record_fft=True, record_var=True), seed=s)
    sw = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, init="poisson", N=500, scale_rates=False,
              local_kernel_sigma=0.005, site_capacity=1)
    betas = np.linspace(0, 3, 64)
    for bi in (0, 9, 18, 27, 36, 45, 54, 63):
        for r in range(6):
            c[f"c2_T20_b{bi:02d}_r{r}"] = dict(ps=dict(sw, beta=float(betas[bi])),
                                               profile=dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.35),
                                               run=dict(T=20.0, obs_dt=0.1, record_fft=True, record_var=True),
                                               seed=10_000 * bi + r)
    for beta in (0.5, 2.5):
        for r in range(3):
            c[f"c3_T40_b{beta}_r{r}"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.05, rate_active=5, beta=beta, init="fixed",
                                                     N=900, scale_rates=False, local_kernel_sigma=0.005, site_capacity=1),
                                             run=dict(T=40.0, obs_dt=1.0, record_fft=True, record_var=True),
                                             seed=303_000 + int(beta * 10) * 10 + r)
    for N, beta in ((290, 1.0), (770, 2.4)):
        for r in range(3):
            c[f"c4_T10_N{N}_r{r}"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.005, rate_active=10, beta=beta, init="poisson",
                                                  N=N, scale_rates=False, local_kernel_sigma=0.02, site_capacity=1),
                                          profile=dict(L=1000, N=N, frac_plus=0.75, decay_plus=0.2),
                                          run=dict(T=10.0, obs_dt=0.1, record_fft=False, record_var=False),
                                          seed=404_000 + N + r)
    for beta in (0.6, 2.1):
        for r in range(2):
            c[f"g0_T20_b{beta}_r{r}"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.002, rate_active=5, beta=beta, init="poisson",
                                                     N=500, scale_rates=False, local_kernel_sigma=0.0, site_capacity=1),
                                             profile=dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.35),
                                             run=dict(T=20.0, obs_dt=0.1, record_fft=False, record_var=True),
                                             seed=505_000 + int(beta * 10) * 10 + r)
    return c


def _one(item):
    name, spec = item
    import gen_golden as G
    from common import digest_out

    PS = G.import_reference()
    rng = CountingRNG(np.random.default_rng(spec["seed"]))
    t0 = time.time()
    ps = G.build_ps(PS, spec, rng)
    out = ps.run(**spec["run"])
    d = digest_out(out)
    n_obs = d["n_obs"]
    rec = dict(ps=spec["ps"], run=spec["run"], seed=spec["seed"], profile=spec.get("profile"),
               n=int(out["pos_list"][0].size), n_events=int(rng.n_exp), n_uniform=int(rng.n_random), seconds=round(time.time() - t0, 1), **d)
    rows = dict(m_local_first=out["m_local_list"][0].copy(), m_local_last=out["m_local_list"][n_obs - 1].copy(),
                pos_last=out["pos_list"][n_obs - 1].astype(np.int32), m_global=out["m_global"].copy())
    if out["fft_amp_list"] is not None:
        rows["fft_amp_head"] = out["fft_amp_list"][:, :8].copy()
    print(f"{name}: n={rec['n']} events={rec['n_events']} n_obs={n_obs} {rec['seconds']} s", flush=True)
    return name, rec, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    items = [(k, v) for k, v in full_cases().items() if not a.names or k in a.names]
    with mp.Pool(a.procs) as pool:
        res = pool.map(_one, items, chunksize=1)
    meta = {name: rec for name, rec, _ in res}
    total = sum(r["n_events"] for r in meta.values())
    json.dump(dict(numpy=np.__version__, scipy=__import__("scipy").__version__, total_events=total, runs=meta),
              open(os.path.join(OUT, "full_length.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "full_length_rows.npz"),
                        **{f"{name}/{k}": v for name, _, rows in res for k, v in rows.items()})
    print(f"{len(meta)} runs, {total} reference events pinned")


if __name__ == "__main__":
    main()
