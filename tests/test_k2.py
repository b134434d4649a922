"""K2 (sublattice-parallel kernel).  K2 has no reference counterpart (the reference cannot run large
lattices); it is the discrete-time sublattice version of the same model, so parity is
  * bit-exact GPU == oracle restatement of the same update rule (include/aps_k2_model.h),
  * bit-exact multi-slab (ghost zones, world_size 2/3, gloo) == single slab,
  * statistical agreement with the exact Gillespie chain (oracle K1) in the dt -> 0 regime,
  * conservation / exclusion invariants at any size."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from aps_b200 import capi
from aps_b200.sublattice import SublatticeLattice, TILE, fixed_point_taps

HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_lattice(L, **kw):
    from oracle_k2 import OracleK2Backend
    return SublatticeLattice(L, backend=OracleK2Backend(), **kw)


PARAMS = dict(D=0.3, lam=3.0, beta=1.2, dt=0.01)


def test_invariants_and_walls():
    for sigma in [None, 2.5]:
        lat = oracle_lattice(2 * TILE, sigma_sites=sigma, seed=3, **PARAMS)
        lat.init_random(0.5, 0.7)
        s0 = lat.state.numpy().copy()
        lat.run(25)
        s1 = lat.state.numpy()
        assert set(np.unique(s1)) <= {0, 1, 2}
        assert (s1 != 0).sum() == (s0 != 0).sum()                      # particles conserved (reflecting walls)
        assert not np.array_equal(s0, s1)
        if sigma is None:                                               # running sum(sigma) stays exact
            assert int(lat.msum[0][0]) == int((s1 == 1).sum()) - int((s1 == 2).sum())


def test_drift_piles_particles_at_the_right_wall():
    lat = oracle_lattice(TILE, sigma_sites=None, seed=5, D=0.05, lam=5.0, beta=0.0, dt=0.02)
    lat.init_random(0.3, 1.0)
    lat.run(400)                                  # T = 8: '+' particles drift ~ lam*T*(1-rho)/2 ~ 15 sites
    rp, rm = lat.profile(TILE // 16)
    tot = rp + rm
    assert tot[-1] > 0.45 and tot[0] < 0.2 and abs(tot[100:400].mean() - 0.3) < 0.02


def _slab_worker(rank, world, port, q, sigma, passes):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    lat = oracle_lattice(6 * TILE, sigma_sites=sigma, seed=11, **PARAMS)
    lat.init_random(0.5, 0.6)
    lat.refresh_every = 20                       # force several ghost refreshes
    lat.run_passes(passes)
    full = lat.gather_state()
    rp, rm = lat.profile(24)
    q.put((rank, full, rp, rm, lat.n_particles, int(lat.msum[0][0])))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("sigma", [3.0, None])
@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_is_bit_identical_to_single_slab(world, sigma):
    """sigma = None: global magnetisation — every rank counts the flips of its own segments only and the increments are
    all-reduced after every pass, so all ranks use the same m as the single-slab run."""
    passes = 70
    single = oracle_lattice(6 * TILE, sigma_sites=sigma, seed=11, **PARAMS)
    single.init_random(0.5, 0.6)
    single.run_passes(passes)
    want = single.state.numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7 * world + (100 if sigma is None else 0)
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, q, sigma, passes)) for r in range(world)]
    [p.start() for p in procs]
    got = [q.get(timeout=600) for _ in range(world)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rp1, rm1 = single.profile(24)
    for rank, full, rp, rm, n, msum in got:
        assert np.array_equal(full, want), f"rank {rank}: slab run differs from the single-slab run"
        assert np.array_equal(rp, rp1) and np.array_equal(rm, rm1) and n == single.n_particles
        if sigma is None:
            assert msum == int(single.msum[0][0]) == int((want == 1).sum()) - int((want == 2).sum())


def test_small_dt_limit_matches_exact_gillespie():
    """Global-field mode, L = 8192, N ~ 4100: magnetisation relaxation and the coarse density profile after
    T = 1.5 against the exact chain (oracle K1, Philox).  Both are single realisations of a self-averaging
    lattice; tolerance = 5 combined standard errors of the bin counts + 2 % (dt bias at dt = 0.004)."""
    from aps_b200.batch import make_params
    from common import HostRun
    from oracle import oracle
    L, T = TILE, 1.5
    D, lam, beta, dt = 0.2, 2.0, 0.6, 0.004
    lat = oracle_lattice(L, sigma_sites=None, seed=21, D=D, lam=lam, beta=beta, dt=dt)
    lat.init_random(0.5, 0.9)
    s0 = lat.state.numpy().copy()
    lat.run(int(round(T / dt)))
    s1 = lat.state.numpy()
    pos0 = np.nonzero(s0)[0].astype(np.int32)
    sg0 = np.where(s0[pos0] == 1, 1, -1).astype(np.int8)
    n = len(pos0)
    times = np.array([0.0, T])                  # the run itself goes on to T + 0.01 so that row 1 is reached
    hr = HostRun(L, n, 2, [n], pos0, sg0, [beta], times, None, seeds=np.array([77], np.uint64), record=1)
    P = make_params(L, 1, -1, D, lam, T + 0.01)
    assert oracle.load().aps_oracle_run(P, hr.batch, 1, 1) == 0
    assert hr.n_obs[0] == 2
    cp, cm = hr.obs_cp[0, 1].astype(int), hr.obs_cm[0, 1].astype(int)
    m_exact = (cp.sum() - cm.sum()) / n
    m_k2 = ((s1 == 1).sum() - (s1 == 2).sum()) / n
    m0 = (sg0.sum()) / n
    assert abs(m0) > 0.7 and abs(m_exact) < 0.45                       # it did relax
    assert abs(m_k2 - m_exact) < 5 * 2 / np.sqrt(n) + 0.02
    nb = 16
    k2_tot = ((s1 != 0).reshape(nb, -1)).sum(1)
    ex_tot = (cp + cm).reshape(nb, -1).sum(1)
    se = np.sqrt(k2_tot + ex_tot)
    assert (np.abs(k2_tot - ex_tot) < 5 * se + 0.02 * ex_tot).all(), (k2_tot, ex_tot)


# ---------------------------------------------------------------- GPU ----------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("sigma", [None, 0.4, 5.0, 40.0])
@pytest.mark.parametrize("tiles", [1, 3])
def test_gpu_pass_equals_oracle_bitwise(sigma, tiles):
    L = tiles * TILE
    kw = dict(sigma_sites=sigma, seed=99, **PARAMS)
    g = SublatticeLattice(L, **kw)
    o = oracle_lattice(L, **kw)
    g.init_random(0.55, 0.65)
    o.init_random(0.55, 0.65)
    assert np.array_equal(g.state.cpu().numpy(), o.state.numpy())       # init kernel == oracle init
    assert g.n_particles == o.n_particles
    for chunk in [1, 2, 7, 20]:
        g.run_passes(chunk); o.run_passes(chunk)
        assert np.array_equal(g.state.cpu().numpy(), o.state.numpy()), (sigma, tiles, chunk)
    if sigma is None:
        assert int(g.msum[0][0]) == int(o.msum[0][0])
    gp, gm = g.profile(40)
    op, om = o.profile(40)
    assert np.array_equal(gp, op) and np.array_equal(gm, om)


@pytest.mark.gpu
@pytest.mark.parametrize("cap", [1, 2, 32])
@pytest.mark.parametrize("sigma", [0.1, 5.0])
def test_gpu_local_field_capacities_do_not_change_results(cap, sigma):
    """Local-field K2 stashes the trials of a tile and evaluates the flip candidates in a dense pass; trials beyond
    the stash (and candidates beyond the list) are replayed inline.  Any capacity must give the oracle's bits
    (cap = 1, 2: almost everything goes through the inline path and the candidate list overflows)."""
    lib = capi.load()
    kw = dict(sigma_sites=sigma, seed=7, D=0.3, lam=3.0, beta=1.2, dt=0.02)     # ~4.4 trials per half and pass
    L = 2 * TILE
    try:
        lib.aps_debug_set_k2_stash_cap(cap)
        g = SublatticeLattice(L, **kw)
        o = oracle_lattice(L, **kw)
        g.init_random(0.6, 0.5); o.init_random(0.6, 0.5)
        for chunk in [1, 9]:
            g.run_passes(chunk); o.run_passes(chunk)
            assert np.array_equal(g.state.cpu().numpy(), o.state.numpy()), (cap, sigma, chunk)
    finally:
        lib.aps_debug_set_k2_stash_cap(0)


@pytest.mark.gpu
def test_gpu_large_lattice_invariants():
    """2^24 sites (16 MiB per buffer): conservation, exclusion and a flat interior profile at beta = 0."""
    L = 1 << 24
    lat = SublatticeLattice(L, D=0.02, lam=5.0, beta=0.0, dt=0.01, sigma_sites=5.0, seed=1)
    lat.init_random(0.5, 0.5)
    n0 = lat.n_particles
    lat.run(50)
    s = lat.state
    assert int((s != 0).sum()) == n0 and int(s.max()) <= 2
    rp, rm = lat.profile(64)
    assert np.allclose((rp + rm)[4:-4], 0.5, atol=0.01)


@pytest.mark.gpu
@pytest.mark.parametrize("sigma", [None, 5.0])
def test_config5_size_lattice_against_the_exact_chain(sigma):
    """Size-independent property at BASELINE config 5's FULL size (2^26 sites, 3.4e7 particles): the relaxation of the
    magnetisation and the mean drift per particle are local quantities, so one huge K2 lattice must reproduce what the EXACT
    Gillespie chain (K1, reference-pinned; for the local field with the reference's double-precision taps) gives on an ensemble
    of 768 lattices of 8192 sites with the same parameters (tools/k2_dt_bias.py -> tests/golden/k2_dt_bias.json: D = 0.2,
    lambda = 2, beta = 0.6, 90 % '+', T = 1.5; global magnetisation and Gaussian local field of 5 sites).
    Tolerance: 3 standard errors of the exact-chain ensemble + 0.002 (the paired dt-bias estimates of the table stay below it)."""
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "k2_dt_bias.json")))
    rows = gold["modes"]["global" if sigma is None else f"local_sigma_{sigma:g}"]
    exact = rows[0]
    assert exact["method"].startswith("exact")
    L, dt = 1 << 26, 0.005
    lat = SublatticeLattice(L, D=gold["D"], lam=gold["lam"], beta=gold["beta"], dt=dt, sigma_sites=sigma, seed=11, single_rank=True)
    lat.init_random(0.5, 0.9)
    n0 = lat.n_particles
    idx = torch.arange(L, device="cuda", dtype=torch.int64)
    x0 = int((idx * (lat.state != 0)).sum()) / n0
    lat.run(int(round(gold["T"] / dt)))
    lat.check()
    s = lat.state
    assert int((s != 0).sum()) == n0 and int(s.max()) <= 2
    m = (int((s == 1).sum()) - int((s == 2).sum())) / n0
    disp = int((idx * (s != 0)).sum()) / n0 - x0
    assert abs(m - exact["m_mean"]) <= 3 * exact["m_se"] + 0.002, (m, exact["m_mean"])
    se_d = max(r["displacement_bias_se"] for r in rows)
    assert abs(disp - exact["displacement_sites"]) <= 3 * se_d + 0.002, (disp, exact["displacement_sites"])


@pytest.mark.gpu
def test_config5_profiles_against_the_reference_pde():
    """BASELINE config 5: K2 profiles against the reference's IMEXPDE (tests/golden/stat_pde_config5.npz, produced by
    the unmodified IMEX_PDE_solver_class.py in tools/gen_golden.py pde_fixture) under the hydrodynamic scaling
    x = site/L, gamma = D/L^2, lam = lambda/L, in the regime where both describe the same dynamics (SURVEY R6:
    only '+' advected, dilute so that exclusion is negligible, global magnetisation, bump away from the walls).
    8 independent lattices of 2^17 sites (~2100 particles in total).  Tolerances: magnetisation 0.04 (3 sigma of the
    particle noise + O(rho) exclusion), drift of the centre of mass 4 %, width 10 %, coarse profile L1 distance 0.2."""
    import json
    from common import GOLDEN
    z = np.load(os.path.join(GOLDEN, "stat_pde_config5.npz"))
    c = json.loads(str(z["meta"]))
    L, T, dt = c["L_lattice"], c["T"], 0.02
    x = (np.arange(L) + 0.5) / L
    rho0 = c["peak_density"] * np.exp(-np.abs(x - 0.5) / 0.05)
    tot_counts = np.zeros(L // 1024)
    n_plus = n_minus = 0
    xs = []
    for seed in range(8):
        rng = np.random.default_rng(1000 + seed)
        occ = rng.random(L) < rho0
        state = np.where(occ, np.where(rng.random(L) < c["frac_plus"], 1, 2), 0).astype(np.uint8)
        lat = SublatticeLattice(L, D=c["D"], lam=c["lam"], beta=c["beta"], dt=dt, sigma_sites=None, seed=seed)
        lat.set_state(state)
        lat.run(int(round(T / dt)))
        s = lat.state.cpu().numpy()
        assert (s != 0).sum() == occ.sum()
        n_plus += int((s == 1).sum()); n_minus += int((s == 2).sum())
        xs.append(x[s != 0])
        tot_counts += (s != 0).reshape(-1, 1024).sum(1)
    xs = np.concatenate(xs)
    pde_tot = z["rho_p"] + z["rho_m"]
    xp = np.arange(1000) / 1000.0
    mean_pde = (pde_tot * xp).sum() / pde_tot.sum()
    std_pde = np.sqrt((pde_tot * xp ** 2).sum() / pde_tot.sum() - mean_pde ** 2)
    m_pde = float(z["m_series"][-1])
    m_k2 = (n_plus - n_minus) / (n_plus + n_minus)
    assert abs(m_k2 - m_pde) < 0.04, (m_k2, m_pde)
    assert abs((xs.mean() - 0.5) - (mean_pde - 0.5)) < 0.04 * (mean_pde - 0.5), (xs.mean(), mean_pde)
    assert abs(xs.std() - std_pde) < 0.10 * std_pde, (xs.std(), std_pde)
    nb = 20
    h_k2 = np.histogram(xs, bins=nb, range=(0, 1))[0] / len(xs)
    h_pde = pde_tot.reshape(nb, -1).sum(1) / pde_tot.sum()
    assert np.abs(h_k2 - h_pde).sum() < 0.2, np.abs(h_k2 - h_pde).sum()


# ---------------------------------------------------------------- persistent launches, multi-GPU ----------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("sigma", [None, 5.0])
def test_gpu_persistent_launch_equals_per_pass_launches(sigma):
    """One cooperative launch with grid barriers between the passes == one launch per pass (and == the oracle, by the
    tests above): same bits for any chunking of the passes, in both field modes."""
    kw = dict(sigma_sites=sigma, seed=5, **PARAMS)
    a = SublatticeLattice(5 * TILE, persistent=True, **kw)
    b = SublatticeLattice(5 * TILE, persistent=False, **kw)         # host loop over aps_k2_pass_device launches
    assert a.persistent and not b.persistent
    a.init_random(0.5, 0.7); b.init_random(0.5, 0.7)
    for chunk in [1, 2, 13, 40]:
        a.run_passes(chunk); b.run_passes(chunk)
        assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy()), (sigma, chunk)
        if sigma is None:
            assert int(a.msum[0][0]) == int(b.msum[0][0])
    a.check()
    assert a.passes_done == b.passes_done == 56


def _gpu_slab_worker(rank, world, port, q, sigma, passes):
    try:
        sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        capi.check(capi.load().aps_set_device(rank), "aps_set_device")
        torch.distributed.init_process_group("nccl", rank=rank, world_size=world)
        kw = dict(sigma_sites=sigma, seed=11, **PARAMS)
        lat = SublatticeLattice(4 * world * TILE, **kw)
        lat.init_random(0.5, 0.6)
        lat.refresh_every = 6                     # several in-kernel ghost refreshes (the strict bound allows >= 150 passes)
        for chunk in [7, passes - 7]:
            lat.run_passes(chunk)
        lat.check()
        full = lat.gather_state()
        rp, rm = lat.profile(32)
        one = SublatticeLattice(4 * world * TILE, single_rank=True, **kw)
        one.init_random(0.5, 0.6)
        one.run_passes(passes)
        want = one.state.cpu().numpy()
        rp1, rm1 = one.profile(32)
        ok = bool(np.array_equal(full, want)) and bool(np.array_equal(rp, rp1)) and bool(np.array_equal(rm, rm1))
        if sigma is None:
            ok = ok and int(lat.msum[0][0]) == int(one.msum[0][0]) == int((want == 1).sum()) - int((want == 2).sum())
        lat.close()
        q.put((rank, ok, ""))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    except Exception as exc:                      # report instead of hanging the parent on q.get
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


@pytest.mark.gpu
@pytest.mark.parametrize("sigma", [5.0, None])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_in_kernel_exchange_is_bit_identical_to_single_gpu(world, sigma):
    """Slab decomposition over `world` GPUs of one box (one process per GPU): ghost refresh and, for sigma = None, the per-pass
    sum(sigma) exchange happen INSIDE the persistent kernel through CUDA-IPC peer memory (NVLink).  The gathered lattice, the
    coarse profile and the running magnetisation must equal the single-GPU run bit for bit.  Skipped on boxes with fewer GPUs."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    passes = 45
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 11 * world + (200 if sigma is None else 0)
    procs = [ctx.Process(target=_gpu_slab_worker, args=(r, world, port, q, sigma, passes)) for r in range(world)]
    [p.start() for p in procs]
    got = [q.get(timeout=600) for _ in range(world)]
    [p.join(120) for p in procs]
    for rank, ok, err in got:
        assert ok, f"rank {rank}: slab run differs from the single-GPU run {err}"
