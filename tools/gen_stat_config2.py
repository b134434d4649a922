#!/usr/bin/env python
"""Reference ensemble for the native-mode statistical parity test on a BASELINE configuration:
N_RUNS runs of the UNMODIFIED reference at the config-2 parameters (sweep_beta.py:837-878; T=20, obs_dt=0.1) for
beta in {0.5, 2.0}, seeded numpy Generators through the `rng=` seam.  Stored per beta (tests/golden/stat_config2.npz):
  * mbar[r]          time average of m_global over the second half of the observation rows, one value per run
                     (the sample of the KS test / of the device histogram);
  * per-site mean and variance over runs of the second-half time averages of rho_plus, rho_minus and m_local
    (the 3-standard-error profile checks).
Build container only (needs /root/reference).  ~30 CPU-minutes, spread over the host cores."""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
N_RUNS = 200
BETAS = [0.5, 2.0]
PS = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, init="poisson", N=500, scale_rates=False,
          local_kernel_sigma=0.005, site_capacity=1)
PROFILE_PLUS = dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.35)      # rho0_plus  = make_exp_gradient(decay .35)[0]  (:859-868)
PROFILE_MINUS_DECAY = 0.2                                                # rho0_minus = make_exp_gradient(decay .2)[1]: flat (:869-878)
RUN = dict(T=20.0, obs_dt=0.1, record_fft=False, record_var=False)


def one(task):
    bi, r = task
    import gen_golden as G
    PSc = G.import_reference()
    spec = dict(ps=dict(PS, beta=BETAS[bi]), profile=PROFILE_PLUS, run=RUN, seed=7_000_000 + 1000 * bi + r)
    ps = G.build_ps(PSc, spec, np.random.default_rng(spec["seed"]))
    out = ps.run(**RUN)
    M = len(out["times_obs"])
    assert all(p is not None for p in out["pos_list"])
    h = slice(M // 2, M)
    return (bi, r, float(out["m_global"][h].mean()), out["rho_p_list"][h].mean(0), out["rho_m_list"][h].mean(0),
            out["m_local_list"][h].mean(0), int(out["pos_list"][0].size))


def main():
    tasks = [(bi, r) for bi in range(len(BETAS)) for r in range(N_RUNS)]
    with mp.Pool(int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()) as pool:
        res = pool.map(one, tasks, chunksize=4)
    save = dict(meta=np.array(json.dumps(dict(ps=PS, profile_plus=PROFILE_PLUS, minus_decay=PROFILE_MINUS_DECAY, run=RUN,
                                              n_runs=N_RUNS, betas=BETAS, numpy=np.__version__))))
    for bi, beta in enumerate(BETAS):
        rows = sorted([x for x in res if x[0] == bi], key=lambda x: x[1])
        save[f"b{bi}_mbar"] = np.array([x[2] for x in rows])
        save[f"b{bi}_n"] = np.array([x[6] for x in rows])
        for k, name in [(3, "rho_p"), (4, "rho_m"), (5, "m_local")]:
            a = np.array([x[k] for x in rows])
            save[f"b{bi}_{name}_mean"] = a.mean(0)
            save[f"b{bi}_{name}_var"] = a.var(0, ddof=1)
        print(f"beta={beta}: <m> = {save[f'b{bi}_mbar'].mean():.4f} +- {save[f'b{bi}_mbar'].std(ddof=1) / np.sqrt(N_RUNS):.4f}")
    np.savez_compressed(os.path.join(OUT, "stat_config2.npz"), **save)


if __name__ == "__main__":
    main()
