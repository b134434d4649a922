"""Per-run structure analyses (aps_b200.structure) against the outputs of the reference's own functions
(PARTICLE_solver_BIOLOGY_local_structure.py:55-103,195-265), recorded in tests/golden/structure.json by
tools/gen_golden.py from the two `reducers_*` reference runs.
CPU part: the batched torch functions fed with the reference's recorded |FFT| rows (first 32 modes) and density rows.
GPU part: the same quantities from a device replay of the recorded trajectory (K1 -> K4 expansion -> cuFFT).
Tolerance 1e-9 relative (FFT implementation and summation order differ); thresholds sit midway between samples."""
import json
import os

import numpy as np
import pytest
import torch

from aps_b200 import structure as st
from common import GOLDEN, load_case

WANT = json.load(open(os.path.join(GOLDEN, "structure.json")))
TAGS = ["b0", "b2"]


def _nan_eq(a, b, rel=1e-9):
    a, b = float(a), float(b)
    return (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, rel=rel, abs=1e-12)


def check_series(amp, total, times, tag):
    w = WANT[tag]
    for k, thr in w["thresholds"].items():
        assert _nan_eq(st.time_to_pattern(amp, times, threshold=thr, k=int(k))[0], w["time_to_pattern"][k])
    assert np.isnan(float(st.time_to_pattern(amp, times, threshold=1e9, k=2)[0]))
    np.testing.assert_allclose(st.lowk_variance_time(amp, k_cut=25)[0].cpu().numpy(), w["lowk_variance_time"], rtol=1e-9)
    assert _nan_eq(st.temporal_autocorrelation(total, lag=1)[0], w["autocorr_lag1"])
    assert _nan_eq(st.temporal_autocorrelation(total, lag=3)[0], w["autocorr_lag3"])
    assert _nan_eq(st.extract_growth_rate(amp, times, k=1, t_min=0.5, t_max=2.5, amp_min=1e-4)[0], w["growth_rate"], rel=1e-7)
    assert _nan_eq(st.extract_growth_rate(amp, times, k=3, t_min=0.0, t_max=None, amp_min=1e-4)[0], w["growth_rate_k3"], rel=1e-7)
    assert np.isnan(float(st.extract_growth_rate(amp, times, k=1, t_min=2.85, amp_min=1e-4)[0]))


@pytest.mark.parametrize("tag", TAGS)
def test_series_analyses_on_reference_rows(tag):
    c = load_case(f"reducers_{tag}")
    amp = torch.from_numpy(c["fft_amp_head"])[None]
    total = torch.from_numpy(c["total_list"])[None]
    check_series(amp, total, c["times_obs"], tag)
    w = WANT[tag]
    assert np.array_equal(st.cluster_size_distribution(c["total_list"][-1], 0.5 * c["total_list"][-1].max()), w["clusters"])
    assert float(st.spectral_entropy(np.array(w["fft_mean_head"]), k_max=25)) == pytest.approx(w["spectral_entropy"], rel=1e-12)


def test_ensemble_time_to_pattern_and_batching():
    amps = torch.stack([torch.from_numpy(load_case(f"reducers_{t}")["fft_amp_head"]) for t in TAGS])
    times = load_case("reducers_b0")["times_obs"]
    e = WANT["ensemble"]
    ttp = st.time_to_pattern(amps, times, threshold=e["threshold"], k=2)
    mean, se = st.ensemble_time_to_pattern(ttp)
    assert mean == pytest.approx(e["mean"], rel=1e-12) and se == pytest.approx(e["se"], rel=1e-12)
    assert all(np.isnan(v) for v in st.ensemble_time_to_pattern(torch.tensor([float("nan")] * 3, dtype=torch.float64)))
    assert np.array_equal(st.cluster_size_distribution(np.zeros(5), 0.1), np.zeros(0, int))
    assert np.array_equal(st.cluster_size_distribution(np.ones(5), 0.1), [5])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_device_structure_analyses_match_the_reference(tag):
    from aps_b200.capi import APS_REC_COUNTS, APS_REC_MLOCAL, APS_REC_POS
    from aps_b200.engine import ReplicaBatch
    c = load_case(f"reducers_{tag}")
    m, w = c["meta"], WANT[tag]
    rb = ReplicaBatch(L=m["L"], K=m["K"], radius=m["radius"], weights=c["weights"], D=m["rate_diffusion"],
                      lam=m["rate_active"], T=m["run"]["T"], times_obs=c["times_obs"], betas=[m["ps"]["beta"]],
                      n=[m["n"]], pos0=c["pos0"][None], sigma0=c["sigma0"][None], dx=m["dx"],
                      record=APS_REC_COUNTS | APS_REC_POS | APS_REC_MLOCAL)
    d = torch.from_numpy(c["draws"]).cuda()
    rb.run_replay(d, torch.tensor([0, len(c["draws"])], dtype=torch.int64, device="cuda"))
    amp, total, var = st.fft_amplitudes(rb)
    check_series(amp, total, c["times_obs"], tag)
    obs = st.structure_observables(rb, start_fraction=0.5, k_max=None, amp=amp, var=var)
    for k in ["var_mean", "low_k_power", "m_local_var", "lowk_variance"]:
        assert float(obs[k][0]) == pytest.approx(w[k], rel=1e-9), k
    assert float(obs["var_std"][0]) == pytest.approx(w["var_std"], abs=1e-12)
    assert int(obs["dominant_k"][0]) == w["dominant_k"]
    np.testing.assert_allclose(obs["fft_mean"][0, :40].cpu().numpy(), w["fft_mean_head"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(obs["fft_std"][0, :40].cpu().numpy(), w["fft_std_head"], rtol=1e-8, atol=1e-9)
    assert float(st.spectral_entropy(obs["fft_mean"], k_max=25)[0]) == pytest.approx(w["spectral_entropy"], rel=1e-9)
    assert float(st.spectral_entropy(obs["fft_mean"])[0]) == pytest.approx(w["spectral_entropy_full"], rel=1e-9)
    assert float(st.mode_competition_ratio(obs["fft_mean"])[0]) == pytest.approx(w["mode_competition"], rel=1e-9)
    last = total[0, -1].cpu().numpy()
    assert np.array_equal(st.cluster_size_distribution(last, 0.5 * last.max()), w["clusters"])
