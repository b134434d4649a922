"""Native (Philox) mode against the reference, statistically (BASELINE north_star): ensemble density
profiles within 3 standard errors and a two-sample KS test on the magnetisation distribution.

Reference side: tests/golden/stat_ensemble.npz = 300 runs per beta of the UNMODIFIED reference with
seeded numpy Generators (tools/gen_golden.py stat_fixture).  Our side: 2000 replicas per beta through
launcher.run_ensemble — on CPU with the oracle-backed shard, on the GPU with the real kernels (the two
are bit-identical, tests/test_launcher.py).

Tolerances (written here on purpose):
  * per site z = |mean_ours - mean_ref| / sqrt(se_ours^2 + se_ref^2): at most 1 % of the 200 (site,species)
    values per beta may exceed 3, none may exceed 4.5 (with 200 comparisons ~0.5 exceed 3 by chance);
  * lattice-aggregated chi^2 / dof < 1.5;
  * KS two-sample p-value > 0.01 on the per-replica time-averaged m_global (broad at beta = 2).
"""
import json
import os

import numpy as np
import pytest
from scipy import stats

from aps_b200 import capi, launcher as la
from aps_b200.capi import APS_REC_COUNTS
from common import GOLDEN

R_OURS = 2000


def _ours(beta, ensemble_cls):
    z = np.load(os.path.join(GOLDEN, "stat_ensemble.npz"))
    meta = json.loads(str(z["meta"]))
    ps = dict(meta["ps"], flip_rate_fn=None, periodic=False, anchor_positions=None, crowding_suppresses_rates=False)
    spec = la.build_beta_sweep_spec([beta], R_OURS, ps, {}, meta["run"], base_seed=424242)
    spec.record = APS_REC_COUNTS
    spec.point_of = np.arange(R_OURS)            # one "point" per replica -> per-replica time-averaged profiles
    res = la.run_ensemble(spec, want_profiles=True, ensemble_cls=ensemble_cls)
    assert (res.status == 0).all()
    return z, res


def _check(bi, beta, ensemble_cls):
    z, res = _ours(beta, ensemble_cls)
    n_ref = len(z[f"b{bi}_mbar"])
    prof = res.profiles                                   # [R][4][L]; rows M//2..M like the fixture
    for q, key in [(0, "rho_p"), (1, "rho_m")]:
        ours_mean, ours_var = prof[:, q].mean(0), prof[:, q].var(0, ddof=1)
        ref_mean, ref_var = z[f"b{bi}_{key}_mean"], z[f"b{bi}_{key}_var"]
        se = np.sqrt(ours_var / R_OURS + ref_var / n_ref) + 1e-12
        zz = np.abs(ours_mean - ref_mean) / se
        assert zz.max() < 4.5, (key, beta, zz.max())
        assert (zz > 3).mean() <= 0.01 + 1e-9, (key, beta, (zz > 3).sum())
        assert (zz ** 2).mean() < 1.5, (key, beta, (zz ** 2).mean())
    # magnetisation: time-average of m_global over the second half = (sum_l rho_p - rho_m)/(sum_l total)
    m_ours = (prof[:, 0].sum(1) - prof[:, 1].sum(1)) / (prof[:, 0].sum(1) + prof[:, 1].sum(1))
    ks = stats.ks_2samp(m_ours, z[f"b{bi}_mbar"])
    assert ks.pvalue > 0.01, (beta, ks)
    se_m = np.sqrt(m_ours.var(ddof=1) / R_OURS + z[f"b{bi}_mbar"].var(ddof=1) / n_ref)
    assert abs(m_ours.mean() - z[f"b{bi}_mbar"].mean()) < 3 * se_m
    return m_ours


@pytest.mark.parametrize("bi,beta", [(0, 0.5), (1, 2.0)])
def test_oracle_native_mode_matches_reference_statistics(bi, beta):
    from oracle_ensemble import OracleEnsemble
    m = _check(bi, beta, OracleEnsemble)
    if beta == 2.0:
        assert m.std() > 0.2                       # ordering has started: broad magnetisation distribution


@pytest.mark.gpu
@pytest.mark.parametrize("bi,beta", [(0, 0.5), (1, 2.0)])
def test_gpu_native_mode_matches_reference_statistics(bi, beta):
    _check(bi, beta, None)


@pytest.mark.gpu
def test_mean_field_fixed_point_with_global_magnetisation():
    """Known-answer check the drivers themselves plot (sweep_beta.py:232-254): with the GLOBAL magnetisation
    (local_kernel_sigma = 0) and flip rate exp(-beta*sigma*m) the stationary magnetisation solves m = tanh(beta*m).
    256 replicas of n = 400 particles, beta = 1.5 (m* = 0.8586), no hops: |<m>| over the second half of the run within
    0.03 of m* (finite-size fluctuations ~ 1/sqrt(n) average out over replicas and time); beta = 0.5: |m| ~ 0."""
    from scipy.optimize import brentq
    ps = dict(L=1000, xlim=1, rate_diffusion=0.0, rate_active=0.0, flip_rate_fn=None, init="fixed", N=400, scale_rates=False,
              local_kernel_sigma=0.0, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False)
    run = dict(T=30.0, obs_dt=0.5)
    spec = la.build_beta_sweep_spec([0.5, 1.5], 256, ps, {}, run, base_seed=77)
    res = la.run_ensemble(spec, want_profiles=False)
    m = res.reducers[:, capi.APS_RED_M_MEAN].reshape(2, 256)
    m_star = brentq(lambda x: x - np.tanh(1.5 * x), 0.1, 1.0)
    assert abs(np.abs(m[1]).mean() - m_star) < 0.03, (np.abs(m[1]).mean(), m_star)
    assert np.abs(m[0]).mean() < 0.08                   # paramagnetic side: fluctuations only


# ---- BASELINE config 2 (sweep_beta.py:837-878, T = 20, obs_dt = 0.1), with the device-side histogram -------------
R_C2 = 1024


def _ks_from_histograms(h_a, h_b):
    """Two-sample KS statistic and asymptotic p-value from two histograms on the same bins."""
    na, nb = h_a.sum(), h_b.sum()
    d = np.abs(np.cumsum(h_a) / na - np.cumsum(h_b) / nb).max()
    return d, stats.kstwo.sf(d, int(round(na * nb / (na + nb))))


@pytest.mark.gpu
def test_config2_native_mode_matches_reference_statistics_with_device_histogram():
    """Reference side: tests/golden/stat_config2.npz = 200 runs per beta of the UNMODIFIED reference at the config-2
    parameters (tools/gen_stat_config2.py).  Our side: 1024 native-mode replicas per beta, one launch.
    Tolerances (stated here on purpose):
      * per site and per field (rho_plus, rho_minus, m_local time-averaged over the second half of the rows),
        z = |mean_ours - mean_ref| / sqrt(var_ours (1/R + 1/n_ref)): at most 1.5 % of the 3000 values per beta above 3
        (0.27 % expected), none above 5.5, mean z^2 < 1.4;
      * the magnetisation distribution: device histogram (256 bins on [-1, 1], `aps_m_histogram_device`) against the
        reference sample binned identically: KS p > 0.01; exact two-sample KS on the raw values: p > 0.01;
        means within 3 standard errors."""
    import torch
    from aps_b200.capi import APS_REC_MLOCAL
    z = np.load(os.path.join(GOLDEN, "stat_config2.npz"))
    meta = json.loads(str(z["meta"]))
    betas = meta["betas"]
    ps = dict(meta["ps"], flip_rate_fn=None, periodic=False, anchor_positions=None, crowding_suppresses_rates=False)
    pp = meta["profile_plus"]
    g = la.make_exp_gradient(L=pp["L"], N=pp["N"], frac_plus=pp["frac_plus"], decay_length=pp["decay_plus"], anchor_positions=None)
    g2 = la.make_exp_gradient(L=pp["L"], N=pp["N"], frac_plus=pp["frac_plus"], decay_length=meta["minus_decay"], anchor_positions=None)
    run = dict(T=meta["run"]["T"], obs_dt=meta["run"]["obs_dt"])
    spec = la.build_beta_sweep_spec(betas, R_C2, ps, dict(rho0_plus=g[0], rho0_minus=g2[1]), run, base_seed=9_000_000)
    spec.record = APS_REC_COUNTS | APS_REC_MLOCAL
    ens = la.DeviceEnsemble(spec, 0, len(spec.betas))
    ens.init_particles()
    ens.rb.run_philox()
    rb = ens.rb
    M = rb.M
    per_rep = rb.profile_sums(1)                                          # [2R][4][L]: time-averaged rho_p, rho_m per replica
    m_loc = rb.obs_m_local[:, M // 2:].mean(dim=1)                         # [2R][L]
    hist, mbar = rb.m_histogram(len(betas), ens.point_of)
    torch.cuda.synchronize()
    assert (rb.status == 0).all() and (rb.n_obs == M).all()
    fields = dict(rho_p=per_rep[:, 0].cpu().numpy(), rho_m=per_rep[:, 1].cpu().numpy(), m_local=m_loc.cpu().numpy())
    hist, mbar = hist.cpu().numpy(), mbar.cpu().numpy()
    edges = np.linspace(-1.0, 1.0, 257)
    for bi, beta in enumerate(betas):
        sl = slice(bi * R_C2, (bi + 1) * R_C2)
        n_ref = len(z[f"b{bi}_mbar"])
        for key, a in fields.items():
            mean_o, var_o = a[sl].mean(0), a[sl].var(0, ddof=1)
            se = np.sqrt(var_o * (1.0 / R_C2 + 1.0 / n_ref)) + 1e-12
            zz = np.abs(mean_o - z[f"b{bi}_{key}_mean"]) / se
            assert zz.max() < 5.5, (key, beta, zz.max(), int(zz.argmax()))
            assert (zz > 3).mean() <= 0.015, (key, beta, int((zz > 3).sum()))
            assert (zz ** 2).mean() < 1.4, (key, beta, (zz ** 2).mean())
        # the device histogram is the histogram of the per-replica values
        want = np.histogram(mbar[sl], bins=edges)[0]
        assert hist[bi].sum() == R_C2 and np.abs(hist[bi] - want).sum() <= 2          # a value on a bin edge may round either way
        ref_hist = np.histogram(z[f"b{bi}_mbar"], bins=edges)[0]
        d, p = _ks_from_histograms(hist[bi], ref_hist)
        assert p > 0.01, (beta, d, p)
        ks = stats.ks_2samp(mbar[sl], z[f"b{bi}_mbar"])
        assert ks.pvalue > 0.01, (beta, ks)
        se_m = np.sqrt(mbar[sl].var(ddof=1) / R_C2 + z[f"b{bi}_mbar"].var(ddof=1) / n_ref)
        assert abs(mbar[sl].mean() - z[f"b{bi}_mbar"].mean()) < 3 * se_m, (beta, mbar[sl].mean(), z[f"b{bi}_mbar"].mean())
