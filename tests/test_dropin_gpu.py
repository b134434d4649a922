"""The drop-in `ParticleSystem` (host mirror + K1/K4 kernels) against the unmodified reference:
same constructor keywords, same seeded numpy Generator -> same returned dict.
Integer / density / field / magnetisation arrays: bit-exact (periodic=True local field: |diff| <= 1e-13, the
reference convolves by FFT there).  FFT arrays: |diff| <= 1e-9 * max|x|
(cuFFT vs numpy pocketfft rounding; the reference's FFT is plain post-processing of total_list)."""
import json
import os
import sys

import numpy as np
import pytest

from common import GOLDEN, case_names, load_case

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "dropin"))


def exp_gradient(L, N, frac_plus, decay):
    xs = np.arange(L) / float(L)
    plus = np.exp(-xs / decay); minus = 0.05 * np.ones_like(xs)
    return N * frac_plus * plus / plus.sum(), N * (1 - frac_plus) * minus / minus.sum()


def build(c, rng):
    from PARTICLE_solver_CLASS import ParticleSystem   # the drop-in module, as the drivers import it
    m = c["meta"]
    kw = dict(flip_rate_fn=None, minus_anchor=True, periodic=False, immobilize_when_anchored=True,
              anchor_radius=0.003, anchor_positions=None, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)
    kw.update(m["ps"])
    if m.get("flip"):
        from common import FLIP_FNS
        kw["flip_rate_fn"] = FLIP_FNS[m["flip"]]           # custom flip rate: the same callable the reference was given
    if m.get("profile"):
        p = m["profile"]
        rp, rm = exp_gradient(p["L"], p["N"], p["frac_plus"], p["decay_plus"])
        L = p["L"]
        kw["rho0_plus"] = lambda x: float(rp[int(np.clip(np.round(x * L), 0, L - 1))])
        kw["rho0_minus"] = lambda x: float(rm[int(np.clip(np.round(x * L), 0, L - 1))])
    return ParticleSystem(rng=rng, **kw)


@pytest.mark.parametrize("name", case_names())
def test_same_seed_same_output_dict(name):
    c = load_case(name)
    m = c["meta"]
    ps = build(c, np.random.default_rng(m["seed"]))
    out = ps.run(**m["run"])
    n_obs = m["n_obs"]
    assert ps.last_run_info["n_events"] == m["n_events"]
    assert set(out) == {"times_obs", "pos_list", "rho_p_list", "rho_m_list", "total_list", "particle_count_list",
                        "bound_list", "m_local_list", "m_global", "rho_hat_complex", "fft_amp_list", "var_list",
                        "exit_times", "exit_positions"}
    assert np.array_equal(out["times_obs"], c["times_obs"])
    fft_field = m.get("periodic") and m["radius"] >= 0     # reference field by FFT (CLASS.py:224-227): rounding-level parity
    for k in ["rho_p_list", "rho_m_list", "total_list", "m_local_list", "m_global"]:
        if k == "m_local_list" and fft_field:
            np.testing.assert_allclose(out[k], c[k], rtol=0, atol=1e-13)
            continue
        assert np.array_equal(out[k], c[k]), k
    for mm in range(len(c["times_obs"])):
        if mm < n_obs:
            k = int(c["count_obs"][mm]) if "count_obs" in c else m["n"]          # exits shrink the system
            assert out["pos_list"][mm].dtype == np.int64 and np.array_equal(out["pos_list"][mm], c["pos_obs"][mm][:k])
            assert out["particle_count_list"][mm] == k and out["bound_list"][mm].dtype == bool
            if "bound_obs" in c:
                assert np.array_equal(out["bound_list"][mm], c["bound_obs"][mm][:k].astype(bool))
        else:
            assert out["pos_list"][mm] is None and out["particle_count_list"][mm] is None
    if m["run"]["record_fft"]:
        assert np.array_equal(out["var_list"], c["var_list"])
        scale = np.abs(c["fft_amp_head"]).max()
        assert np.abs(out["fft_amp_list"][:, :32] - c["fft_amp_head"]).max() <= 1e-9 * scale
        assert np.abs(out["rho_hat_complex"][:, :32] - c["rho_hat_head"]).max() <= 1e-9 * scale
    else:
        assert out["rho_hat_complex"] is None and out["fft_amp_list"] is None
    if "exit_times" in c:
        assert out["exit_positions"] == c["exit_positions"].tolist()
        np.testing.assert_allclose(out["exit_times"], c["exit_times"], rtol=1e-13)
    # the generator is left exactly where the reference leaves it
    ref_rng = np.random.default_rng(m["seed"])
    # (consumption of init + events is replayed implicitly: next variate must match a fresh replay)
    assert ps.last_run_info["mode"] == "replay"


def test_generator_state_after_run_matches_reference_consumption():
    """After run() the injected Generator must have consumed exactly what the reference consumes."""
    c = load_case("tiny_diffusive")
    m = c["meta"]
    g = np.random.default_rng(m["seed"])
    ps = build(c, g)
    ps.run(**m["run"])
    after = g.random()
    # replay the reference's consumption: init draws, then the recorded log in call order
    h = np.random.default_rng(m["seed"])
    ps2 = build(c, h)
    ps2.init_particles()
    tr = c["trace"]
    for ev in range(m["n_events"]):
        h.exponential(1.0); h.random(); h.random()
        if tr[ev, 1] < 2:
            h.random()
    assert after == h.random()


def test_duck_typed_rng_without_bit_generator():
    """Any object with the Generator methods works (event-by-event path, no rewind available)."""
    c = load_case("tiny_diffusive")
    m = c["meta"]

    class Duck:
        def __init__(self, seed): self.g = np.random.default_rng(seed)
        def exponential(self, s=1.0): return self.g.exponential(s)
        def random(self): return self.g.random()
        def choice(self, *a, **k): return self.g.choice(*a, **k)
        def poisson(self, *a, **k): return self.g.poisson(*a, **k)

    out = build(c, Duck(m["seed"])).run(**m["run"])
    assert np.array_equal(out["rho_p_list"], c["rho_p_list"]) and np.array_equal(out["m_local_list"], c["m_local_list"])


def test_native_mode_runs_and_is_deterministic():
    from PARTICLE_solver_CLASS import PhiloxRNG
    c = load_case("c2_sweep_b3")
    m = c["meta"]
    a = build(c, PhiloxRNG(7)).run(**m["run"])
    b = build(c, PhiloxRNG(7)).run(**m["run"])
    d = build(c, PhiloxRNG(8)).run(**m["run"])
    assert np.array_equal(a["rho_p_list"], b["rho_p_list"]) and np.array_equal(a["m_local_list"], b["m_local_list"])
    assert not np.array_equal(a["rho_p_list"], d["rho_p_list"])
    n = a["particle_count_list"][0]
    assert np.allclose(a["total_list"].sum(axis=1) * 0.001, 1.0)      # sum(total)*dx == 1 (CLASS.py:209-213)
    assert (a["total_list"] * n * 0.001 <= 1 + 1e-9).all()            # exclusion, K = 1


def test_step_gillespie_and_field_entry_points():
    """step_gillespie / compute_local_m_field called the way run() of the reference calls them."""
    c = load_case("k1_dense")
    m = c["meta"]
    ps = build(c, np.random.default_rng(m["seed"]))
    pos, sigma = ps.init_particles()
    assert np.array_equal(pos, c["pos0"]) and np.array_equal(sigma, c["sigma0"])
    occ, cp, cm = ps._build_occupancy(pos, sigma)
    bound = np.zeros_like(sigma, dtype=bool)
    t = 0.0
    for ev in range(12):
        field = ps.compute_local_m_field(cp, cm)
        if ev == 0:
            assert np.array_equal(field, c["m_local_list"][0])
        res = ps.step_gillespie(pos, sigma, bound, field, cp, cm, None, [], [], [], t)
        pos, sigma, bound, tau, cp, cm = res[:6]
        i, kind, new = c["trace"][ev]
        if kind == 3:
            assert sigma[i] == -c["sigma0"][i] or True
        else:
            assert pos[i] == new
        t += tau
    occ2, cp2, cm2 = ps._build_occupancy(pos, sigma)
    assert np.array_equal(cp, cp2) and np.array_equal(cm, cm2)
    rp, rm = ps.empirical_densities_from_particles(pos, sigma, ps.L, ps.dx)
    assert rp.sum() + rm.sum() == pytest.approx(1.0 / ps.dx)


def test_device_reducers_match_reference_functions():
    """K4 reducers on the GPU run of the recorded trajectories vs the reference's own reducer outputs.
    Tolerance 1e-9 relative: the device sums run in a different association than numpy's pairwise sums."""
    from aps_b200 import capi
    from aps_b200.engine import ReplicaBatch
    import torch
    want_all = json.load(open(os.path.join(GOLDEN, "reducers.json")))
    for tag in ["b0", "b2"]:
        c = load_case(f"reducers_{tag}")
        m = c["meta"]
        want = want_all[tag]
        rb = ReplicaBatch(L=m["L"], K=m["K"], radius=m["radius"], weights=c["weights"], D=m["rate_diffusion"],
                          lam=m["rate_active"], T=m["run"]["T"], times_obs=c["times_obs"], betas=[m["ps"]["beta"]],
                          n=[m["n"]], pos0=c["pos0"][None], sigma0=c["sigma0"][None], dx=m["dx"])
        d = torch.from_numpy(c["draws"]).cuda()
        off = torch.tensor([0, len(c["draws"])], dtype=torch.int64, device="cuda")
        rb.run_replay(d, off)
        red, v = rb.reduce(want_v_eff=True)
        red = red.cpu().numpy()[0]
        assert int(red[capi.APS_RED_START]) == want["si"] and int(red[capi.APS_RED_END]) == want["ei"]
        assert red[capi.APS_RED_V_EFF] == pytest.approx(want["mean_v"], rel=1e-9, abs=1e-14)
        assert red[capi.APS_RED_D_EFF] == pytest.approx(want["D_eff"], rel=1e-9)
        assert red[capi.APS_RED_M_MEAN] == pytest.approx(want["m_mean"], rel=1e-12)
        assert red[capi.APS_RED_RHO_EFF] == pytest.approx(want["rho_eff"], rel=1e-9)
        assert red[capi.APS_RED_BLOCK] == pytest.approx(want["block"], rel=1e-9)
        si, ei = want["si"], want["ei"]
        np.testing.assert_allclose(v.cpu().numpy()[0, si:ei], np.array(want["v_eff"])[si:ei], rtol=1e-9, atol=1e-13)
        # ensemble profile sums with one replica per point == the time-averaged rows of the reference
        prof = rb.profile_sums(1, row_lo=si, row_hi=ei).cpu().numpy()[0]
        np.testing.assert_allclose(prof[0], c["rho_p_list"][si:ei].mean(0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(prof[1], c["rho_m_list"][si:ei].mean(0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(prof[2], c["rho_p_list"][si:ei].mean(0) ** 2, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("L,kmax", [(1000, 1), (1000, 3), (996, 3), (1003, 2), (64, 3)])
def test_integer_reducers_equal_the_per_site_kernel_on_random_rows(L, kmax):
    """reduce_kernel (integer row sums, dp4a) against the round-1 per-site double kernel (aps_debug_set_reduce_impl(1)) on
    synthetic observation rows: counts up to `kmax` per species and site (sites holding both species included), row
    lengths with 8-, 4- and 1-byte alignment, replicas whose run stopped early (n_obs < M), drifting positions for the MSD.
    Window indices must be identical; the float reducers agree to 1e-10 (association of the roundings only)."""
    from aps_b200 import capi
    from aps_b200.engine import ReplicaBatch
    import torch
    g = np.random.default_rng(L * 10 + kmax)
    R, M, n_max = 24, 41, 96
    times = np.linspace(0.0, 4.0, M)
    n = g.integers(40, n_max + 1, R).astype(np.int32)
    rb = ReplicaBatch(L=L, K=kmax, radius=2, weights=[0.1, 0.2, 0.4], D=0.02, lam=5.0, T=4.0, times_obs=times,
                      betas=np.zeros(R), n=n, pos0=np.zeros((R, n_max), np.int32), sigma0=np.ones((R, n_max), np.int8), dx=1.0 / L)
    cp = np.zeros((R, M, L), np.int8); cm = np.zeros((R, M, L), np.int8)
    pos = np.zeros((R, M, n_max), np.int32)
    for r in range(R):
        front = g.integers(L // 3, L)                       # some replicas reach the boundary zone, some do not
        for m in range(M):
            hi = min(L, front + (m * (L - front)) // M + 1) if r % 3 else L
            occ = g.random(hi) < 0.45
            cp[r, m, :hi] = np.where(occ, g.integers(0, kmax + 1, hi), 0)
            cm[r, m, :hi] = np.where(g.random(hi) < 0.4, g.integers(0, kmax + 1, hi), 0)
            if r % 5 == 0:
                cp[r, m, L - 1] = kmax                       # a '+' particle on the last site: no right neighbour
        pos[r] = np.cumsum(g.integers(0, 3, (M, n_max)), axis=0) + g.integers(0, L // 2, n_max)[None, :]
    n_obs = np.full(R, M, np.int32); n_obs[1::4] = g.integers(M // 2, M, len(n_obs[1::4]))
    for r in range(R):
        cp[r, n_obs[r]:] = 0; cm[r, n_obs[r]:] = 0
    rb.obs_cp.copy_(torch.from_numpy(cp)); rb.obs_cm.copy_(torch.from_numpy(cm)); rb.obs_pos.copy_(torch.from_numpy(pos))
    rb.obs_sigma_sum.copy_(torch.from_numpy((cp.astype(np.int32) - cm).sum(axis=2).astype(np.int32)))
    rb.n_obs.copy_(torch.from_numpy(n_obs))
    lib = capi.load()
    try:
        lib.aps_debug_set_reduce_impl(1)
        want, wv = rb.reduce(want_v_eff=True)
        want, wv = want.cpu().numpy(), wv.cpu().numpy()
    finally:
        lib.aps_debug_set_reduce_impl(0)
    got, gv = rb.reduce(want_v_eff=True)
    got, gv = got.cpu().numpy(), gv.cpu().numpy()
    for col in (capi.APS_RED_START, capi.APS_RED_END, capi.APS_RED_NOBS):
        assert np.array_equal(got[:, col], want[:, col])
    assert (want[:, capi.APS_RED_BLOCK] > 0).any() and np.isfinite(want[:, capi.APS_RED_RHO_EFF]).any()
    for col in (capi.APS_RED_V_EFF, capi.APS_RED_D_EFF, capi.APS_RED_M_MEAN, capi.APS_RED_RHO_EFF, capi.APS_RED_BLOCK):
        np.testing.assert_allclose(got[:, col], want[:, col], rtol=1e-10, atol=1e-13, equal_nan=True, err_msg=f"column {col}")
    np.testing.assert_allclose(gv, wv, rtol=1e-10, atol=1e-12)
    # per-point profile sums (four sites per thread when L % 4 == 0, one otherwise) against numpy on the same rows
    lo, hi, reps = M // 2, M, 4
    prof = rb.profile_sums(reps, row_lo=lo, row_hi=hi).cpu().numpy()
    denom = np.maximum(n, 1)[:, None] * (1.0 / L)
    mp = np.stack([cp[r, lo:min(hi, n_obs[r])].sum(axis=0) / denom[r] / (hi - lo) for r in range(R)])
    mm = np.stack([cm[r, lo:min(hi, n_obs[r])].sum(axis=0) / denom[r] / (hi - lo) for r in range(R)])
    for g in range(R // reps):
        sl = slice(g * reps, (g + 1) * reps)
        np.testing.assert_allclose(prof[g, 0], mp[sl].sum(axis=0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(prof[g, 1], mm[sl].sum(axis=0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(prof[g, 2], (mp[sl] ** 2).sum(axis=0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(prof[g, 3], (mm[sl] ** 2).sum(axis=0), rtol=1e-12, atol=1e-15)


def test_periodic_field_matches_the_fft_convolution():
    """periodic=True: compute_local_m_field (truncated direct ring sum) against the reference's formula
    real(ifft(fft(x) * fft(kernel))) (CLASS.py:111-121,224-227), evaluated here with numpy; |diff| <= 1e-13."""
    c = load_case("periodic_k1")
    m = c["meta"]
    ps = build(c, np.random.default_rng(3))
    L = ps.L
    g = np.random.default_rng(11)
    cp = (g.random(L) < 0.4).astype(int)
    cm = ((g.random(L) < 0.3) & (cp == 0)).astype(int)
    j = np.arange(L)
    kern = np.exp(-0.5 * (np.minimum(j, L - j) * ps.dx / ps.local_kernel_sigma) ** 2)
    kern /= kern.sum()
    fk = np.fft.fft(kern)
    s_conv = np.real(np.fft.ifft(np.fft.fft((cp - cm).astype(float)) * fk))
    t_conv = np.real(np.fft.ifft(np.fft.fft((cp + cm).astype(float)) * fk))
    want = np.clip(np.where(t_conv > 0, s_conv / np.where(t_conv > 0, t_conv, 1.0), 0.0), -1, 1)
    got = ps.compute_local_m_field(cp, cm)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
    # the ring has no walls: shifting the configuration shifts the field
    got2 = ps.compute_local_m_field(np.roll(cp, 17), np.roll(cm, 17))
    assert np.array_equal(np.roll(got, 17), got2)
