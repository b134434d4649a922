"""Randomised parity sweep: the K1 kernels (through the C ABI, host buffers) against the CPU oracle on seeded random
model parameters — every combination of site capacity, global / local / wide (radius > L) / periodic field, crowding,
anchors with binding / unbinding / exit, ragged replica sizes, native (Philox) and replay mode.
Bit-exact in every output (state, observation rows, field, clock, trace, exit log); no tolerances."""
import numpy as np
import pytest

from aps_b200 import capi
from aps_b200.batch import make_params
from aps_b200.engine import gaussian_weights, periodic_weights
from common import HostRun, assert_same_outputs, run_oracle

pytestmark = pytest.mark.gpu


def random_case(seed, sorted_init=False, lean=False):
    """sorted_init: initial positions in increasing order (what the device Poisson init produces);
    lean: restrict to what the half-image kernel aps_k1_lean.cuh accepts (K = 1, local field with r <= 20, plain model)."""
    g = np.random.default_rng(seed)
    L = int(g.integers(8, 300)) if not lean else int(g.integers(8, 1000))
    K = int(g.choice([1, 1, 2, 3, 5])) if not lean else 1
    kind = g.choice(["global", "local", "wide", "periodic", "periodic_global"]) if not lean else "local"
    flags = 0
    if kind in ("global", "periodic_global"):
        radius, weights = -1, np.zeros(1)
    elif kind == "local" and lean:
        radius, weights = gaussian_weights(float(g.uniform(0.2, min(5.1, max(0.3, (L - 1) / 4.2)))))   # r <= 20 and r < L
    elif kind == "local":
        radius, weights = gaussian_weights(float(g.uniform(0.2, max(0.3, L / 12))))
    elif kind == "wide":
        radius, weights = gaussian_weights(float(g.uniform(L / 5, L / 2.5)))      # radius ~ L .. 1.6 L: repeated reflection
    else:
        radius, weights = periodic_weights(L, 1.0 / L, float(g.uniform(0.3, max(0.4, L / 40))) / L)
    if kind.startswith("periodic"):
        flags |= capi.APS_FLAG_PERIODIC
    if g.random() < 0.3 and not lean:
        flags |= capi.APS_FLAG_CROWDING
    anchors = g.random() < 0.35 and not lean
    k_on = k_off = k_exit = 0.0
    mask = None
    if anchors:
        mask = (g.random(L) < 0.15).astype(np.uint8)
        k_on, k_off, k_exit = float(g.uniform(0.5, 8)), float(g.uniform(0.5, 8)), float(g.choice([0.0, g.uniform(0.5, 6)]))
        flags |= (capi.APS_FLAG_SUPPRESS_FLIP_BOUND if g.random() < 0.5 else 0) | (capi.APS_FLAG_IMMOBILIZE if g.random() < 0.5 else 0)
    D = float(g.choice([0.0, g.uniform(0.05, 2.0)]))
    lam = float(g.uniform(0.2, 5.0))
    R = 4
    n_cap = max(1, min(int(0.8 * K * L), 300))
    ns = [int(g.integers(1, n_cap + 1)) for _ in range(R)]
    n_max = max(ns)
    pos0 = np.zeros((R, n_max), np.int32)
    sigma0 = np.ones((R, n_max), np.int8)
    for r, n in enumerate(ns):
        slots = np.repeat(np.arange(L), K)                      # respects the capacity
        pos0[r, :n] = g.permutation(slots)[:n]
        if sorted_init:
            pos0[r, :n] = np.sort(pos0[r, :n])
        sigma0[r, :n] = g.choice([1, -1], n)
    betas = g.uniform(0.0, 3.0, R)
    M = int(g.integers(3, 12))
    T = 500.0 / (n_max * (2 * D + lam + 2.0))
    times = np.arange(M) * (T / M)
    params = make_params(L, K, radius, D, lam, T, flags, k_on=k_on, k_off=k_off, k_exit=k_exit)
    return dict(L=L, n_max=n_max, M=M, ns=ns, pos0=pos0, sigma0=sigma0, betas=betas, times=times, weights=weights,
                params=params, mask=mask, kind=kind, K=K)


def build(c, seeds=None, draws=None, draw_off=None):
    return HostRun(c["L"], c["n_max"], c["M"], c["ns"], c["pos0"], c["sigma0"], c["betas"], c["times"], c["weights"],
                   seeds=seeds, draws=draws, draw_off=draw_off, trace_cap=4000, anchor_mask=c["mask"],
                   exit_cap=c["n_max"] if c["mask"] is not None else 0)


@pytest.mark.parametrize("seed", range(48))
def test_native_mode_random_parameters(seed):
    lib = capi.load()
    c = random_case(1000 + seed)
    seeds = np.array([seed, 2 ** 33 + seed, 7 * seed + 1, 2 ** 63 + seed], np.uint64)
    gpu = build(c, seeds=seeds)
    capi.check(lib.aps_run_philox_host(c["params"], gpu.batch), "aps_run_philox_host")
    ora = run_oracle(c["params"], build(c, seeds=seeds), mode=1, threads=2)
    assert_same_outputs(gpu, ora)
    assert (gpu.n_events > 0).any(), c["kind"]


@pytest.mark.parametrize("seed", range(24))
def test_replay_mode_random_parameters(seed):
    lib = capi.load()
    c = random_case(5000 + seed)
    g = np.random.default_rng(seed)
    lens = g.integers(40, 1500, len(c["ns"]))                  # some replicas run out of variates mid-run
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    draws = g.random(int(off[-1]))
    gpu = build(c, draws=draws, draw_off=off)
    capi.check(lib.aps_run_replay_host(c["params"], gpu.batch), "aps_run_replay_host")
    ora = run_oracle(c["params"], build(c, draws=draws, draw_off=off), mode=0, threads=2)
    assert_same_outputs(gpu, ora)
    assert set(gpu.status.tolist()) <= {capi.APS_RUN_DONE, capi.APS_RUN_DRAWS_EXHAUSTED, capi.APS_RUN_EMPTY}


@pytest.mark.parametrize("seed", range(32))
def test_half_image_kernel_random_parameters(seed):
    """Sorted initial positions, K = 1, local field with r <= 20: the configurations aps_k1_lean.cuh takes (28 replicas per SM),
    native and replay mode alternating; ragged replica sizes, lattices from 8 to 1000 sites."""
    lib = capi.load()
    c = random_case(9000 + seed, sorted_init=True, lean=True)
    assert c["params"].radius <= 20 and c["params"].radius < c["L"] and c["K"] == 1
    if seed % 2 == 0:
        seeds = np.array([seed, 2 ** 33 + seed, 7 * seed + 1, 2 ** 63 + seed], np.uint64)
        gpu = build(c, seeds=seeds)
        capi.check(lib.aps_run_philox_host(c["params"], gpu.batch), "aps_run_philox_host")
        ora = run_oracle(c["params"], build(c, seeds=seeds), mode=1, threads=2)
    else:
        g = np.random.default_rng(seed)
        lens = g.integers(200, 3000, len(c["ns"]))
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        draws = g.random(int(off[-1]))
        gpu = build(c, draws=draws, draw_off=off)
        capi.check(lib.aps_run_replay_host(c["params"], gpu.batch), "aps_run_replay_host")
        ora = run_oracle(c["params"], build(c, draws=draws, draw_off=off), mode=0, threads=2)
    assert_same_outputs(gpu, ora)
    assert (gpu.n_events > 0).any()


@pytest.mark.parametrize("n", [1, 7, 8, 9, 128, 129, 257, 488, 489, 495, 505, 512, 600, 968, 969, 1000, 1024])
def test_half_image_kernel_particle_count_edges(n):
    """Sorted K = 1 inputs around the sizes where numpy's pairwise-sum tree changes shape (129: two leaves, 257: three,
    489: five leaves -> the half-image kernel hands the replica to the full-size kernel; 969: nine leaves -> the full-size
    kernel hands it to the generic one): all must equal the oracle."""
    lib = capi.load()
    g = np.random.default_rng(n)
    L = 700 if n <= 512 else 1030
    radius, weights = gaussian_weights(3.0)
    R = 3
    ns = [n, max(1, n - 3), n]
    pos0 = np.zeros((R, n), np.int32); sigma0 = np.ones((R, n), np.int8)
    for r, k in enumerate(ns):
        pos0[r, :k] = np.sort(g.choice(L, k, replace=False)); sigma0[r, :k] = g.choice([1, -1], k)
        if r == 1:                       # one replica in random particle order: site-map kernels (lean n > 512, fast otherwise)
            pos0[r, :k] = g.permutation(pos0[r, :k])
    M = 4
    T = 300.0 / (n * 4.0)
    params = make_params(L, 1, radius, 0.3, 2.0, T, 0)
    c = dict(L=L, n_max=n, M=M, ns=ns, pos0=pos0, sigma0=sigma0, betas=np.array([0.5, 1.5, 2.5]), times=np.arange(M) * (T / M),
             weights=weights, params=params, mask=None)
    seeds = np.array([n, n + 1, n + 2], np.uint64)
    gpu = build(c, seeds=seeds)
    capi.check(lib.aps_run_philox_host(params, gpu.batch), "aps_run_philox_host")
    ora = run_oracle(params, build(c, seeds=seeds), mode=1, threads=2)
    assert_same_outputs(gpu, ora)
    assert (gpu.n_events > 20).all()
