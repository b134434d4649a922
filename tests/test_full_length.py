"""Full-length reference pins (tests/golden/full_length.json, written by tools/gen_golden_full.py from the UNMODIFIED
reference): 66 runs of the BASELINE configurations at their full length — config 1 (T=20), config 2 (T=20, 8 betas x 6
seeds), config 3 (T=40), config 4 (T=10, r=80), the global-field point of sweep_beta_2 — 1 072 611 reference events.

A reference run is a pure function of (keywords, seed); the drop-in consumes the same seeded numpy Generator in the
reference's call order, so equality of the event count and of the digests of every returned array pins the whole
trajectory.  Bar: bit-exact (sha256 of the raw arrays) for positions, densities, the local field, m_global and the
variance; the FFT amplitudes (different FFT library on the device) to 1e-9 of their scale.

  * CPU (`-m "not gpu"`): the oracle, driven through the drop-in's own replay driver, on a subset (~1.8e5 events);
  * GPU (`-m gpu`): the CUDA path (K1 + K4 through the C ABI) on all 66 runs.
"""
import json
import os
import sys

import numpy as np
import pytest

from common import GOLDEN, digest_out

sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "dropin"))

PINS = json.load(open(os.path.join(GOLDEN, "full_length.json")))
RUNS = PINS["runs"]
ROWS = np.load(os.path.join(GOLDEN, "full_length_rows.npz"))
DIGEST_KEYS = ["pos", "rho_p_list", "rho_m_list", "total_list", "m_local_list", "m_global", "var_list"]
CPU_SUBSET = ["c1_T20_s1", "c2_T20_b00_r0", "c2_T20_b27_r3", "c2_T20_b63_r5", "c3_T40_b2.5_r0", "c4_T10_N770_r1", "g0_T20_b2.1_r0"]


def test_pins_cover_a_million_reference_events():
    assert PINS["total_events"] == sum(r["n_events"] for r in RUNS.values()) >= 1_000_000
    assert len(RUNS) == 66


def build(rec, rng):
    from PARTICLE_solver_CLASS import ParticleSystem   # the drop-in module, as the drivers import it
    from test_dropin_gpu import exp_gradient
    kw = dict(flip_rate_fn=None, minus_anchor=True, periodic=False, immobilize_when_anchored=True,
              anchor_radius=0.003, anchor_positions=None, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)
    kw.update(rec["ps"])
    if rec.get("profile"):
        p = rec["profile"]
        rp, rm = exp_gradient(p["L"], p["N"], p["frac_plus"], p["decay_plus"])
        L = p["L"]
        kw["rho0_plus"] = lambda x: float(rp[int(np.clip(np.round(x * L), 0, L - 1))])
        kw["rho0_minus"] = lambda x: float(rm[int(np.clip(np.round(x * L), 0, L - 1))])
    return ParticleSystem(rng=rng, **kw)


def check(name, rec, out, n_events):
    assert n_events == rec["n_events"], f"{name}: {n_events} events, the reference made {rec['n_events']}"
    d = digest_out(out)
    assert d["n_obs"] == rec["n_obs"]
    n_obs = d["n_obs"]
    # small rows first: a mismatch here is readable, a digest mismatch is not
    assert np.array_equal(out["m_global"], ROWS[f"{name}/m_global"]), name
    assert np.array_equal(out["pos_list"][n_obs - 1], ROWS[f"{name}/pos_last"]), name
    assert np.array_equal(out["m_local_list"][0], ROWS[f"{name}/m_local_first"]), name
    assert np.array_equal(out["m_local_list"][n_obs - 1], ROWS[f"{name}/m_local_last"]), name
    for k in DIGEST_KEYS:
        if k in rec:
            assert d[k] == rec[k], f"{name}: digest of {k} differs from the reference's"
    if out.get("fft_amp_list") is not None:
        want = ROWS[f"{name}/fft_amp_head"]
        assert np.abs(out["fft_amp_list"][:, :8] - want).max() <= 1e-9 * np.abs(want).max()


def _oracle_out(ps, rec):
    """ParticleSystem.run() with the device batch replaced by the oracle (same replay driver, same aps_batch descriptor)."""
    from test_replay_driver import OracleBackedBatch
    pos, sigma = ps.init_particles()
    times_obs = np.arange(0.0, rec["run"]["T"], rec["run"]["obs_dt"])
    c = dict(meta=dict(L=ps.L, K=ps.K, radius=ps._radius, rate_diffusion=ps.rate_diffusion, rate_active=ps.rate_active,
                       run=rec["run"], ps=rec["ps"], n=int(pos.size)),
             times_obs=times_obs, weights=ps._weights)
    rb = OracleBackedBatch(c, pos, sigma)
    ps._run_replay(rb)
    hr = rb.hr
    n, M, n_obs = int(pos.size), len(times_obs), int(hr.n_obs[0])
    denom = float(max(1, n)) * ps.dx
    rho_p = np.zeros((M, ps.L)); rho_m = np.zeros((M, ps.L)); m_loc = np.zeros((M, ps.L)); m_glob = np.zeros(M); var = np.zeros(M)
    rho_p[:n_obs] = hr.obs_cp[0, :n_obs].astype(np.int64) / denom            # CLASS.py:205-213
    rho_m[:n_obs] = hr.obs_cm[0, :n_obs].astype(np.int64) / denom
    total = rho_p + rho_m
    m_loc[:n_obs] = hr.obs_m_local[0, :n_obs]
    m_glob[:n_obs] = hr.obs_sigma_sum[0, :n_obs] / float(n)
    for m in range(n_obs):
        var[m] = np.var(total[m])                                            # :501,529
    out = dict(pos_list=[hr.obs_pos[0, m, :n].astype(np.int64) if m < n_obs else None for m in range(M)],
               particle_count_list=[n if m < n_obs else None for m in range(M)],
               rho_p_list=rho_p, rho_m_list=rho_m, total_list=total, m_local_list=m_loc, m_global=m_glob,
               var_list=var if (rec["run"]["record_fft"] and rec["run"]["record_var"]) else None, fft_amp_list=None)
    if rec["run"]["record_var"] and not rec["run"]["record_fft"]:
        out["var_list"] = np.zeros(M)                                        # allocated but never filled (:499-507,527-535)
    return out, int(hr.n_events[0])


@pytest.mark.parametrize("name", CPU_SUBSET)
def test_oracle_reproduces_full_length_reference_runs(name):
    rec = RUNS[name]
    ps = build(rec, np.random.default_rng(rec["seed"]))
    out, n_ev = _oracle_out(ps, rec)
    check(name, rec, out, n_ev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(RUNS))
def test_cuda_path_reproduces_full_length_reference_runs(name):
    rec = RUNS[name]
    ps = build(rec, np.random.default_rng(rec["seed"]))
    out = ps.run(**rec["run"])
    assert ps.last_run_info["mode"] == "replay"
    check(name, rec, out, ps.last_run_info["n_events"])
