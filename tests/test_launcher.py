"""Launcher: sharding, collectives (gloo, world_size 2, on CPU) and the reference-shaped sweep API."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from aps_b200 import launcher as la

HERE = os.path.dirname(os.path.abspath(__file__))

PS = dict(L=64, xlim=1, rate_diffusion=0.4, rate_active=3, flip_rate_fn=None, init="poisson", N=30, scale_rates=False,
          local_kernel_sigma=0.03, minus_anchor=True, periodic=False, anchor_positions=None, site_capacity=1,
          crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)
RUN = dict(T=3.0, obs_dt=0.1, record_fft=True, record_var=True)


def init_kwargs():
    g = la.make_exp_gradient(L=64, N=30, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
    return dict(rho0_plus=g[0], rho0_minus=g[1])


def _sweep(betas=(0.0, 1.0, 2.5), runs=3):
    from oracle_ensemble import OracleEnsemble
    return la.sweep_over_betas(list(betas), runs, PS, init_kwargs(), RUN, base_seed=5, ensemble_cls=OracleEnsemble)


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    out = _sweep()
    q.put((rank, {k: v for k, v in out.items() if isinstance(v, np.ndarray)}, out["info"]))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in [0, 1, 7, 64, 4096, 4097]:
        for w in [1, 2, 3, 8]:
            b = [la.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_world2_gloo_equals_single_process():
    single = _sweep()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=300) for _ in range(2)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, arrs, info in got:
        assert info["world"] == 2 and info["shard"] == la.shard_bounds(9, rank, 2)
        for k, v in arrs.items():
            np.testing.assert_array_equal(v, single[k], err_msg=k) if v.dtype.kind in "iu" else \
                np.testing.assert_allclose(v, single[k], rtol=1e-13, atol=1e-15, err_msg=k)


def test_reference_shaped_outputs():
    out = _sweep(betas=(0.5, 2.0), runs=4)
    for k in ["beta_values", "means", "stds", "ses", "D_means", "D_ses", "m_means", "m_stds", "m_ses", "rho_means",
              "rho_ses", "block_means", "block_ses", "ps_kwargs", "outs"]:      # sweep_beta.py:952-970
        assert k in out, k
    assert out["means"].shape == (2,) and out["n_events"].shape == (2, 4)
    assert (out["status"] == 0).all() and (out["n_events"] > 10).all()
    assert out["m_means"][1] != out["m_means"][0]
    assert out["rho_plus_profile_mean"].shape == (2, 64)
    # sum over sites of the time-averaged total density * dx == 1 for every replica (CLASS.py:209-213)
    tot = (out["rho_plus_profile_mean"] + out["rho_minus_profile_mean"]).sum(1) * (1.0 / 64)
    np.testing.assert_allclose(tot, 1.0, rtol=1e-12)


def test_make_exp_gradient_matches_driver_fixture():
    from common import load_case
    c = load_case("c2_sweep_b0")
    g = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
    assert np.array_equal(g[2], c["rho0_plus"]) and np.array_equal(g[3], c["rho0_minus"])
    assert g[0](0.25) == c["rho0_plus"][250]


@pytest.mark.gpu
def test_gpu_launcher_matches_oracle_ensemble():
    """End to end on the GPU (device init + K1 + K4 + gather) vs the oracle-backed ensemble:
    event counts and windows exact, reducers to 1e-9 (different summation association)."""
    from oracle_ensemble import OracleEnsemble
    ik = init_kwargs()
    a = la.sweep_over_betas([0.0, 1.0, 2.5], 3, PS, ik, RUN, base_seed=5)
    b = la.sweep_over_betas([0.0, 1.0, 2.5], 3, PS, ik, RUN, base_seed=5, ensemble_cls=OracleEnsemble)
    assert np.array_equal(a["n_events"], b["n_events"]) and np.array_equal(a["status"], b["status"])
    for k in ["means", "ses", "D_means", "m_means", "rho_means", "block_means"]:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(a["rho_plus_profile_mean"], b["rho_plus_profile_mean"], rtol=1e-12, atol=1e-15)
    # fixed init with per-point N (double-sweep style)
    ps = dict(PS, init="fixed", N=20)
    a = la.sweep_over_betas([0.3, 1.7], 4, ps, {}, RUN, base_seed=9)
    b = la.sweep_over_betas([0.3, 1.7], 4, ps, {}, RUN, base_seed=9, ensemble_cls=OracleEnsemble)
    assert np.array_equal(a["n_events"], b["n_events"])
    np.testing.assert_allclose(a["means"], b["means"], rtol=1e-9, atol=1e-13)
    # periodic ring (generic K1 kernel, truncated ring kernel)
    ps = dict(PS, init="fixed", N=20, periodic=True)
    a = la.sweep_over_betas([0.5, 2.0], 3, ps, {}, RUN, base_seed=11)
    b = la.sweep_over_betas([0.5, 2.0], 3, ps, {}, RUN, base_seed=11, ensemble_cls=OracleEnsemble)
    assert np.array_equal(a["n_events"], b["n_events"]) and a["n_events"].min() > 0
    np.testing.assert_allclose(a["means"], b["means"], rtol=1e-9, atol=1e-13)


@pytest.mark.gpu
def test_gpu_init_matches_oracle_init_bitwise():
    from oracle_ensemble import OracleEnsemble
    for init, K in [("poisson", 1), ("poisson", 2), ("fixed", 1), ("fixed", 3)]:
        ps = dict(PS, init=init, site_capacity=K, N=40)
        spec = la.build_beta_sweep_spec([0.5, 1.5], 5, ps, init_kwargs(), RUN, base_seed=3)
        dev = la.DeviceEnsemble(spec, 0, 10)
        dev.init_particles()
        seeds, pos0, sg0, n = OracleEnsemble(spec, 0, 10).init_states()
        assert np.array_equal(dev.n.cpu().numpy(), n)
        for r in range(10):
            assert np.array_equal(dev.pos0[r, :n[r]].cpu().numpy(), pos0[r, :n[r]])
            assert np.array_equal(dev.sigma0[r, :n[r]].cpu().numpy(), sg0[r, :n[r]])
        occ = np.zeros(64, int)
        np.add.at(occ, pos0[0, :n[0]], 1)
        assert occ.max() <= K


def test_oracle_init_statistics_match_reference_init():
    """Device-side init (restated in the oracle) vs the reference's numpy init: mean particle number and
    mean '+' occupancy profile agree within 4 standard errors (different RNG, same law)."""
    from oracle_ensemble import OracleEnsemble
    sys.path.insert(0, os.path.join(HERE, "..", "dropin"))
    from PARTICLE_solver_CLASS import ParticleSystem
    R = 400
    for init, K in [("poisson", 1), ("poisson", 2), ("fixed", 1), ("fixed", 2)]:
        ps = dict(PS, init=init, site_capacity=K, N=40)
        spec = la.build_beta_sweep_spec([1.0], R, ps, init_kwargs(), RUN, base_seed=11)
        seeds, pos0, sg0, n = OracleEnsemble(spec, 0, R).init_states()
        ours_n = n.astype(float)
        ours_plus = np.zeros((R, 64))
        for r in range(R):
            np.add.at(ours_plus[r], pos0[r, :n[r]][sg0[r, :n[r]] == 1], 1)
        ref_n, ref_plus = np.zeros(R), np.zeros((R, 64))
        g = np.random.default_rng(123)
        for r in range(R):
            psys = ParticleSystem(beta=1.0, rng=g, **ps, **init_kwargs())
            p, s = psys.init_particles()
            ref_n[r] = p.size
            np.add.at(ref_plus[r], p[s == 1], 1)
        se = np.sqrt(ours_n.var(ddof=1) / R + ref_n.var(ddof=1) / R) + 1e-12
        assert abs(ours_n.mean() - ref_n.mean()) <= 4 * se, (init, K)
        se_prof = np.sqrt(ours_plus.var(0, ddof=1) / R + ref_plus.var(0, ddof=1) / R) + 1e-9
        z = np.abs(ours_plus.mean(0) - ref_plus.mean(0)) / se_prof
        assert (z < 4.5).all(), (init, K, z.max())


@pytest.mark.gpu
def test_structure_observables_match_the_reference_function():
    """local_structure.py:55-103 restated with numpy on the downloaded run vs the device version.
    Tolerance 1e-9 relative (cuFFT vs pocketfft, different summation order)."""
    ps = dict(PS, init="fixed", N=40, local_kernel_sigma=0.03)
    run = dict(T=4.0, obs_dt=0.25)
    res = la.sweep_betas_for_structures([0.5, 2.0], 3, ps, {}, run, start_fraction=0.5, k_max=None, base_seed=4)
    assert set(res) == {0.5, 2.0}
    # recompute replica by replica with numpy from the raw observation rows
    from aps_b200.capi import APS_REC_COUNTS, APS_REC_MLOCAL, APS_REC_POS
    spec = la.build_beta_sweep_spec([0.5, 2.0], 3, ps, {}, run, base_seed=4)
    spec.record = APS_REC_COUNTS | APS_REC_POS | APS_REC_MLOCAL
    ens = la.DeviceEnsemble(spec, 0, 6)
    ens.init_particles(); ens.rb.run_philox()
    cp, cm = ens.rb.obs_cp.cpu().numpy().astype(np.int64), ens.rb.obs_cm.cpu().numpy().astype(np.int64)
    ml = ens.rb.obs_m_local.cpu().numpy()
    n = ens.n.cpu().numpy()
    M = cp.shape[1]; s = M // 2
    for b, beta in enumerate([0.5, 2.0]):
        var_means, lowk, mlv, fmeans = [], [], [], []
        for j in range(3):
            r = 3 * b + j
            total = cp[r] / (n[r] * (1.0 / 64)) + cm[r] / (n[r] * (1.0 / 64))
            var_ts = np.array([np.var(u) for u in total])
            amp = np.abs(np.fft.fft(total, axis=1))
            fm = amp[s:].mean(axis=0)
            var_means.append(var_ts[s:].mean()); lowk.append(fm[1:25].sum()); mlv.append(np.var(ml[r, s:])); fmeans.append(fm)
        np.testing.assert_allclose(res[beta]["var_mean"], np.mean(var_means), rtol=1e-9)
        np.testing.assert_allclose(res[beta]["low_k_power_mean"], np.mean(lowk), rtol=1e-9)
        np.testing.assert_allclose(res[beta]["m_local_var_mean"], np.mean(mlv), rtol=1e-9)
        np.testing.assert_allclose(res[beta]["fft_mean_mean"], np.mean(fmeans, axis=0), rtol=1e-8, atol=1e-9)


@pytest.mark.gpu
def test_double_sweep_grid():
    ps = dict(PS)
    out = la.double_sweep([10, 30, 50], [0.0, 1.5], 4, ps, RUN, frac_plus=0.75, decay_plus=0.2, base_seed=2)
    assert set(out) == {10, 30, 50, "info"}
    for N in [10, 30, 50]:
        assert out[N]["means"].shape == (2,) and np.isfinite(out[N]["block_means"]).all()
    # denser systems block more (exclusion): blocking probability grows with N
    assert out[50]["block_means"].mean() > out[10]["block_means"].mean()


def test_npz_cache_uses_the_reference_key_names(tmp_path):
    out = _sweep(betas=(0.5, 2.0), runs=3)
    p = la.save_sweep_npz(str(tmp_path / "CHANGES_simulation_out_sweep.npz"), out)
    # the reference's reload path, sweep_beta.py:933-950
    data = np.load(p, allow_pickle=True)
    save_dict = dict(data)
    for k in ["beta_values", "means", "stds", "ses", "D_means", "D_ses", "m_means", "m_stds", "m_ses", "rho_means", "rho_ses",
              "block_means", "block_ses"]:
        assert np.array_equal(save_dict[k], out[k], equal_nan=True), k
    assert save_dict["ps_kwargs"].item()["L"] == 64 and "outs" in save_dict
    assert la.load_sweep_npz(p)["ps_kwargs"]["site_capacity"] == 1


def test_periodic_weights_truncation():
    """Ring kernel taps (CLASS.py:111-121): symmetric, discarded mass <= 1e-22, refuses a kernel wider than the ring."""
    from aps_b200.engine import periodic_weights
    L, dx, s = 200, 1.0 / 200, 0.02
    r, w = periodic_weights(L, dx, s)
    assert w.size == 2 * r + 1 and 2 * r + 1 <= L and np.array_equal(w, w[::-1])
    j = np.arange(L)
    k = np.exp(-0.5 * (np.minimum(j, L - j) * dx / s) ** 2); k /= k.sum()
    assert w[r] == k[0] and w[r + 3] == k[3] and abs(1.0 - w.sum()) < 1e-15
    assert k[r + 1:L - r].sum() <= 1e-22 and (r == 0 or k[r:L - r + 1].sum() > 1e-22)
    with pytest.raises(NotImplementedError):
        periodic_weights(64, 1.0 / 64, 0.5)


def test_balanced_order_gives_every_rank_every_sweep_point():
    """Strided assignment: with replicas listed beta-major, each rank's contiguous block of the permuted spec holds
    the same number of replicas of every beta, and the permutation is a bijection that carries seeds with replicas."""
    nb, reps, world = 6, 8, 4
    spec = la.build_beta_sweep_spec(np.linspace(0, 3, nb), reps, dict(L=64, rate_diffusion=0.1, rate_active=1.0, N=10,
                                                                         scale_rates=False, init="fixed"), {}, dict(T=1.0))
    order = la.balanced_order(len(spec.betas), world)
    assert sorted(order.tolist()) == list(range(nb * reps))
    ps = la.permute_spec(spec, order)
    assert np.array_equal(ps.seeds, spec.seeds[order]) and np.array_equal(ps.betas, spec.betas[order])
    for r in range(world):
        lo, hi = la.shard_bounds(len(spec.betas), r, world)
        assert np.array_equal(order[lo:hi], np.arange(r, nb * reps, world))
        assert np.array_equal(np.bincount(ps.point_of[lo:hi], minlength=nb), np.full(nb, reps // world))
    assert np.array_equal(la.balanced_order(5, 1), np.arange(5))


@pytest.mark.gpu
def test_sweep_over_sigmas_and_raw_structure_series(tmp_path):
    """sweep_beta_2.py:1030-1075 (results keyed by sigma, npz key names) and the light per-run `raw` entries of
    sweep_betas_for_structures that feed the driver's time-series analyses."""
    ps = dict(PS, init="fixed", N=30)
    res = la.sweep_over_sigmas([0.0, 0.03], [0.5, 2.0], 3, ps, {}, RUN, base_seed=3, save_dir=str(tmp_path))
    assert list(res) == [0.0, 0.03]
    for si, sigma in enumerate(res):
        # every sigma of a reproducible sweep gets its own streams: base_seed + si * 1_000_003 * max(10 000, runs per beta)
        one = la.sweep_over_betas([0.5, 2.0], 3, dict(ps, local_kernel_sigma=sigma), {}, RUN, base_seed=3 + si * 1_000_003 * 10_000,
                                  want_profiles=False)
        assert np.array_equal(res[sigma]["v_mean"], one["means"]) and np.array_equal(res[sigma]["D_se"], one["D_ses"])
        z = np.load(tmp_path / f"v_eff_vs_beta_sigma_{sigma:.4g}.npz", allow_pickle=True)
        assert set(z.files) == {"beta", "v_mean", "v_se", "D_mean", "D_se", "ps_kwargs"}
        assert z["ps_kwargs"].item()["local_kernel_sigma"] == sigma
    assert not np.array_equal(res[0.0]["v_mean"], res[0.03]["v_mean"])
    st_res = la.sweep_betas_for_structures([0.5], 4, dict(ps, local_kernel_sigma=0.03), {}, dict(T=4.0, obs_dt=0.25), k_keep=16)
    raw = st_res[0.5]["raw"]
    assert len(raw) == 4 and raw[0]["out"]["fft_amp_list"].shape == (16, 16) and raw[0]["out"]["var_list"].shape == (16,)
    assert np.mean([r["var_mean"] for r in raw]) == pytest.approx(st_res[0.5]["var_mean"], rel=1e-12)
    from aps_b200 import structure as st
    lk = st.lowk_variance_time(torch.from_numpy(raw[0]["out"]["fft_amp_list"])[None], k_cut=10)
    assert lk.shape == (1, 16) and float(lk[0, 8:].mean()) > 0


def test_schedule_order_is_longest_first_inside_balanced_blocks():
    """schedule_order = strided rank assignment, then descending replica_cost inside every rank's block; results are
    returned in the caller's order (covered by test_world2_gloo_equals_single_process)."""
    ik = init_kwargs()
    spec = la.build_beta_sweep_spec(np.linspace(0, 3, 7), 6, PS, ik, RUN)
    cost = la.replica_cost(spec)
    assert cost.shape == (42,) and (cost > 0).all()
    by_beta = cost.reshape(7, 6)[:, 0]
    assert by_beta[-1] > by_beta[0] and by_beta.argmin() not in (0, 6)      # frac_plus = 0.75: minimum at exp(beta) = 3
    for world in (1, 3):
        order = la.schedule_order(spec, world)
        assert sorted(order.tolist()) == list(range(42))
        for r in range(world):
            lo, hi = la.shard_bounds(42, r, world)
            assert sorted(order[lo:hi].tolist()) == list(range(r, 42, world))
            assert (np.diff(cost[order[lo:hi]]) <= 1e-12).all()
    fixed = la.build_beta_sweep_spec([0.0, 2.0], 3, dict(PS, init="fixed", N=20), {}, RUN)
    assert np.allclose(la.replica_cost(fixed), la.replica_cost(fixed)[0])     # m0 = 0: no beta dependence, order untouched
    assert np.array_equal(la.schedule_order(fixed, 1), np.arange(6))
