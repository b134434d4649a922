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
