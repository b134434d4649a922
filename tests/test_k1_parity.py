"""GPU parity tests proper: the K1 kernel, called through the C ABI, against
  (a) the golden vectors of the unmodified reference (trace + every observation row), and
  (b) the CPU oracle on the same inputs — bit-exact in EVERY output including the event clock.
Tolerances: none (integer state, fp64 field and clock compared as bit patterns)."""
import numpy as np
import pytest

from aps_b200 import capi
from common import (HostRun, assert_matches_reference, assert_same_outputs, case_names, hostrun_from_case,
                    load_case, params_from_case, run_oracle)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    lib = capi.load()
    assert lib.aps_device_count() >= 1
    return lib


def run_gpu(lib, params, hr, philox=False):
    fn = lib.aps_run_philox_host if philox else lib.aps_run_replay_host
    capi.check(fn(params, hr.batch), "aps_run_*_host")
    return hr


@pytest.mark.parametrize("threads,use_lut,use_fast", [(0, 1, 1), (0, 1, 2), (64, 0, 0), (128, 1, 0),
                                                      (256, 0, 0), (32, 0, 0), (64, 1, 0)])
@pytest.mark.parametrize("name", case_names())
def test_replay_matches_reference_and_oracle(lib, name, threads, use_lut, use_fast):
    """Every block-size variant, both filter-tap paths (product table / arithmetic) of the generic kernel,
    and the K=1 specialised kernel (taken by the K=1 local-field cases: use_fast=1 capacity-class layout,
    use_fast=2 run-time layout)."""
    c = load_case(name)
    lib.aps_debug_set_k1_threads(threads)
    lib.aps_debug_set_use_lut(use_lut)
    lib.aps_debug_set_use_fast(use_fast)
    try:
        g = run_gpu(lib, params_from_case(c), hostrun_from_case(c))
    finally:
        lib.aps_debug_set_k1_threads(0)
        lib.aps_debug_set_use_lut(1)
        lib.aps_debug_set_use_fast(1)
    assert_matches_reference(c, g)
    o = run_oracle(params_from_case(c), hostrun_from_case(c))
    assert_same_outputs(g, o)
    assert g.n_guard[0] == 0


@pytest.mark.parametrize("name", ["c2_sweep_b3", "tiny_diffusive", "global_sigma0", "crowding", "huge_sigma"])
def test_exact_slow_path_gives_the_same_trajectory(lib, name):
    """Widen the guard band so that most selections go through the serial exact path."""
    c = load_case(name)
    lib.aps_debug_set_guard_scale(1e12)
    try:
        g = run_gpu(lib, params_from_case(c), hostrun_from_case(c))
    finally:
        lib.aps_debug_set_guard_scale(1.0)
    assert g.n_guard[0] > 0
    assert_matches_reference(c, g)


def test_draw_exhaustion_and_resume(lib):
    """Stop on a short log, then resume from the returned state: same result as one run."""
    c = load_case("tiny_diffusive")
    m = c["meta"]
    p = params_from_case(c)
    full = run_gpu(lib, p, hostrun_from_case(c))
    cut = 200
    a = HostRun(m["L"], m["n"], len(c["times_obs"]), [m["n"]], c["pos0"], c["sigma0"], [m["ps"]["beta"]],
                c["times_obs"], c["weights"], draws=c["draws"][:cut], draw_off=[0, cut])
    run_gpu(lib, p, a)
    assert a.status[0] == capi.APS_RUN_DRAWS_EXHAUSTED
    used = int(a.draws_used[0])
    assert 0 < used <= cut
    rest = c["draws"][used:]
    b = HostRun(m["L"], m["n"], len(c["times_obs"]), [m["n"]], a.pos_end, a.sigma_end, [m["ps"]["beta"]],
                c["times_obs"], c["weights"], draws=rest, draw_off=[0, len(rest)],
                t_start=a.t_end, obs_start=a.n_obs, ev_start=a.n_events)
    b.obs_cp[:] = a.obs_cp; b.obs_cm[:] = a.obs_cm; b.obs_pos[:] = a.obs_pos
    b.obs_sigma_sum[:] = a.obs_sigma_sum; b.obs_m_local[:] = a.obs_m_local
    run_gpu(lib, p, b)
    assert b.status[0] == capi.APS_RUN_DONE
    for f in ["obs_cp", "obs_cm", "obs_pos", "obs_sigma_sum", "n_obs", "n_events", "pos_end", "sigma_end"]:
        assert np.array_equal(getattr(b, f), getattr(full, f)), f
    assert np.array_equal(b.obs_m_local.view(np.uint64), full.obs_m_local.view(np.uint64))
    assert b.t_end[0] == full.t_end[0]


def test_batched_replicas_with_ragged_sizes(lib):
    """Several replicas of different n and beta in one launch == the same replicas one by one."""
    names = ["k1_dense", "r0_local", "tiny_diffusive"]
    for name in names:
        c = load_case(name)
        m = c["meta"]
        R, n, nmax = 7, m["n"], m["n"] + 5
        rng = np.random.default_rng(5)
        ns = rng.integers(max(1, n // 2), n + 1, R).astype(np.int32)
        ns[0] = n
        pos0 = np.zeros((R, nmax), np.int32); sg0 = np.ones((R, nmax), np.int8)
        for r in range(R):
            pos0[r, :ns[r]] = c["pos0"][:ns[r]]; sg0[r, :ns[r]] = c["sigma0"][:ns[r]]
        betas = np.linspace(0.0, 3.0, R); betas[0] = m["ps"]["beta"]
        # replica 0 replays the recorded log; the others follow different trajectories, so their log
        # must be valid in every role: values in [0,1) serve as uniforms and as exponential variates
        nd = len(c["draws"])
        draws = np.concatenate([c["draws"]] + [rng.random(nd) for _ in range(R - 1)]); off = np.arange(R + 1) * nd
        mk = lambda: HostRun(m["L"], nmax, len(c["times_obs"]), ns, pos0, sg0, betas, c["times_obs"], c["weights"],
                             draws=draws, draw_off=off)
        g = run_gpu(lib, params_from_case(c), mk())
        o = run_oracle(params_from_case(c), mk(), threads=2)
        assert_same_outputs(g, o)
        assert g.n_events[0] == m["n_events"]


def test_philox_mode_matches_oracle(lib):
    """Native mode: the in-kernel Philox stream and -log(1-u) clock equal the oracle's, bit for bit."""
    for name in ["c2_sweep_b3", "c1_exclusion", "tiny_diffusive", "global_sigma0"]:
        c = load_case(name)
        m = c["meta"]
        R = 6
        seeds = np.array([1, 2, 2**40 + 7, 2**63 + 11, 12345, 0], np.uint64)
        mk = lambda: HostRun(m["L"], m["n"], len(c["times_obs"]), [m["n"]] * R, np.tile(c["pos0"], R),
                             np.tile(c["sigma0"], R), np.linspace(0, 3, R), c["times_obs"], c["weights"], seeds=seeds)
        g = run_gpu(lib, params_from_case(c), mk(), philox=True)
        o = run_oracle(params_from_case(c), mk(), mode=1, threads=2)
        assert_same_outputs(g, o)
        assert (g.status == capi.APS_RUN_DONE).all() and (g.n_events > 0).all()
        assert len(set(g.n_events.tolist())) > 1   # different seeds / betas -> different runs


def test_over_capacity_initial_state_still_matches_oracle(lib):
    """The reference accepts initial states with more than K particles on a site; the product-table
    path cannot index them, so such a replica must take the arithmetic path and still agree."""
    c = load_case("k1_dense")
    m = c["meta"]
    n = m["n"]
    pos0 = c["pos0"].copy(); pos0[:6] = pos0[6]          # seven particles on one site, K = 1
    rng = np.random.default_rng(9)
    draws = rng.random(4000)
    mk = lambda: HostRun(m["L"], n, len(c["times_obs"]), [n], pos0, c["sigma0"], [1.3], c["times_obs"], c["weights"],
                         draws=draws, draw_off=[0, len(draws)], trace_cap=1200)
    g = run_gpu(lib, params_from_case(c), mk())
    o = run_oracle(params_from_case(c), mk())
    assert_same_outputs(g, o)
    assert g.n_events[0] > 50


def test_empty_replica_reports_the_reference_failure(lib):
    c = load_case("k1_dense")
    m = c["meta"]
    hr = HostRun(m["L"], 4, len(c["times_obs"]), [0], np.zeros((1, 4)), np.ones((1, 4)), [1.0], c["times_obs"],
                 c["weights"], draws=c["draws"], draw_off=[0, len(c["draws"])])
    run_gpu(lib, params_from_case(c), hr)
    assert hr.status[0] == capi.APS_RUN_EMPTY and hr.n_events[0] == 0
