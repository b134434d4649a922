#!/usr/bin/env python
"""dt-bias of the sublattice kernel: K2 is a discrete-time (synchronous-sublattice) version of the model, exact
only for dt -> 0.  This tool quotes the bias: R independent lattices of L = 8192 sites (global field, reflecting walls, D=0.2,
lambda=2, beta=0.6, start 90 % '+', density 0.5) are run to T = 1.5 by
  * the EXACT chain: K1 (rejection-free Gillespie, native Philox mode, generic kernel), one replica per lattice;
  * K2 at dt in {0.02, 0.01, 0.005, 0.0025}, one lattice per seed,
and the ensemble means of the magnetisation m(T) and of the mean particle position are compared.  Prints a markdown table
(mean +- standard error, bias = K2 - exact, in units of the combined standard error).
--sigma lists the field modes: 0 = global magnetisation (above), s > 0 = Gaussian local field of s sites — there the exact chain
uses the reference's double-precision scipy taps and K2 its 16-bit fixed-point taps and quantised field (include/aps_k2_model.h),
so the table also bounds the effect of that quantisation (default: both modes)."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200.engine import ReplicaBatch, gaussian_weights
from aps_b200.sublattice import SublatticeLattice, TILE

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=192)
ap.add_argument("--out", default="")
ap.add_argument("--sigma", type=float, nargs="*", default=[0.0, 5.0], help="field modes: 0 = global magnetisation, > 0 = Gaussian local field of that many sites")
a = ap.parse_args()
L, T, D, lam, beta, R = TILE, 1.5, 0.2, 2.0, 0.6, a.replicas
# identical initial states for both methods: K2's Bernoulli init kernel, one seed per lattice
states = []
for s in range(R):
    lat = SublatticeLattice(L, D=D, lam=lam, beta=beta, dt=0.01, sigma_sites=None, seed=1000 + s, single_rank=True)
    lat.init_random(0.5, 0.9)
    states.append(lat.state.cpu().numpy().copy())
states = np.stack(states)
n = (states != 0).sum(1)
n_max = int(n.max())
pos0 = np.zeros((R, n_max), np.int32); sg0 = np.ones((R, n_max), np.int8)
for s in range(R):
    idx = np.nonzero(states[s])[0]
    pos0[s, :len(idx)] = idx; sg0[s, :len(idx)] = np.where(states[s, idx] == 1, 1, -1)
x = (np.arange(L) + 0.5) / L



def run_mode(sigma):
    # ---- exact chain (K1) ----
    times = np.array([0.0, T])
    radius, wts = gaussian_weights(sigma) if sigma > 0 else (-1, np.zeros(1))      # the reference's double-precision taps
    rb = ReplicaBatch(L=L, K=1, radius=radius, weights=wts, D=D, lam=lam, T=T + 0.01, times_obs=times, betas=np.full(R, beta),
                      n=n.astype(np.int32), pos0=pos0, sigma0=sg0, seeds=np.arange(R, dtype=np.uint64) + 77, record=1)
    rb.run_philox()
    torch.cuda.synchronize()
    assert (rb.n_obs == 2).all() and (rb.status == 0).all()
    cp, cm = rb.obs_cp[:, 1].cpu().numpy().astype(int), rb.obs_cm[:, 1].cpu().numpy().astype(int)
    m_exact = (cp.sum(1) - cm.sum(1)) / n
    x_exact = ((cp + cm) * x).sum(1) / n
    rows = [("exact Gillespie chain (K1)", m_exact, x_exact)]
    # ---- K2 at several dt ----
    for dt in [0.02, 0.01, 0.005, 0.0025]:
        mk, xk = [], []
        for s in range(R):
            lat = SublatticeLattice(L, D=D, lam=lam, beta=beta, dt=dt, sigma_sites=(sigma if sigma > 0 else None), seed=5000 + s, single_rank=True)
            lat.set_state(states[s])
            lat.run(int(round(T / dt)))
            st = lat.state.cpu().numpy()
            mk.append(((st == 1).sum() - (st == 2).sum()) / n[s]); xk.append((x * (st != 0)).sum() / n[s])
        rows.append((f"K2, dt = {dt}", np.array(mk), np.array(xk)))
    se = lambda v: v.std(ddof=1) / np.sqrt(len(v))
    # paired statistics: lattice s starts from the same state in every method, so the bias estimate is the mean over lattices of
    # the per-lattice difference (the spread of the initial configurations cancels); positions in lattice sites
    out = ["| method | m(T=1.5) | bias of m (paired) | in SE | mean displacement (sites) | bias (paired, sites) | in SE |", "|---|---|---|---|---|---|---|"]
    x0 = np.array([(x * (states[s] != 0)).sum() / n[s] for s in range(R)])
    res = []
    for name, m, xx in rows:
        dm, dx = m - m_exact, (xx - x_exact) * L
        first = name.startswith("exact")
        disp = (xx - x0) * L
        out.append(f"| {name} | {m.mean():.5f} +- {se(m):.5f} | {'' if first else f'{dm.mean():+.5f} +- {se(dm):.5f}'} | {'' if first else f'{dm.mean() / se(dm):+.1f}'} | "
                   f"{disp.mean():.4f} +- {se(disp):.4f} | {'' if first else f'{dx.mean():+.4f} +- {se(dx):.4f}'} | {'' if first else f'{dx.mean() / se(dx):+.1f}'} |")
        res.append(dict(method=name, m_mean=float(m.mean()), m_se=float(se(m)), m_bias=float(dm.mean()), m_bias_se=float(se(dm)),
                        displacement_sites=float(disp.mean()), displacement_bias_sites=float(dx.mean()), displacement_bias_se=float(se(dx))))
    print(f"\n### {'global magnetisation' if sigma <= 0 else f'Gaussian local field, sigma = {sigma:g} sites (K1: double-precision taps; K2: 16-bit taps, quantised field)'}\n\nR = {R} lattices of L = {L}, ~{int(n.mean())} particles each, m(0) = {float(((states == 1).sum(1) - (states == 2).sum(1)).mean() / n.mean()):.3f}\n")
    print("\n".join(out))
    return res



results = {}
for sg in a.sigma:
    results["global" if sg <= 0 else f"local_sigma_{sg:g}"] = run_mode(sg)
if a.out:
    first = results.get("global") or next(iter(results.values()))
    json.dump(dict(replicas=R, L=L, T=T, D=D, lam=lam, beta=beta, rows=first, modes=results), open(a.out, "w"), indent=1)
