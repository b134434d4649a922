"""Ensemble / sweep launcher: replaces the serial loops of the sweep drivers.

Reference loops replaced (all are `for b in betas: for run in range(n_runs): ParticleSystem(...).run()`):
  sweep_beta_ensemble / sweep_over_betas      PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py:56-117, :828-1028
  (N_part, beta) double sweep                 ..._double_sweep.py:100-152, :665-873
  (sigma, beta) sweep                         ..._sweep_beta_2.py:1030-1075
  sweep_betas_for_structures                  PARTICLE_solver_BIOLOGY_local_structure.py:105-193

Every replica is an independent ParticleSystem (private state and rng, sweep_beta.py:83), so the
(grid point x replica) list is flattened, block-partitioned over the ranks (one process per GPU,
`torch.distributed`), each rank runs its shard in ONE K1 launch (native Philox mode, device-side
initial conditions), reduces every run on the device (K4) and only then communicates:
  * per-replica scalars (8 doubles each)            -> all_gather
  * per-grid-point profile sums [points][4][L] f64  -> all_reduce(sum)   (NCCL over NVLink)
There is no data-path collective inside the time stepping.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import capi
from .capi import APS_REC_COUNTS, APS_REC_POS, APS_RED_N, ApsInitArgs
from .engine import ReplicaBatch, _dev, _stream, gaussian_weights, periodic_weights


def make_exp_gradient(L, N, frac_plus, decay_length, anchor_positions=(0.25, 0.60), anchor_peak_width=0.01,
                      anchor_peak_mass=0.03):
    """Initial-profile helper of the drivers (sweep_beta.py:16-53): exponential '+' profile, flat '-'
    profile with optional Gaussian bumps; returns [rho0_plus(x), rho0_minus(x), rho_plus[], rho_minus[]]."""
    xs = np.arange(L) / float(L)
    plus = np.exp(-xs / decay_length)
    minus = 0.05 * np.ones_like(xs)
    if anchor_positions is not None:
        for a in anchor_positions:
            minus += anchor_peak_mass * np.exp(-0.5 * ((xs - a) / anchor_peak_width) ** 2)
    rho_plus = N * frac_plus * (plus / plus.sum())
    rho_minus = N * (1 - frac_plus) * (minus / minus.sum())

    def rho0_plus(x):
        return float(rho_plus[int(np.clip(np.round(x * L), 0, L - 1))])

    def rho0_minus(x):
        return float(rho_minus[int(np.clip(np.round(x * L), 0, L - 1))])

    # the tabulated profile travels with the callable, so a sweep does not have to call it L times from Python
    rho0_plus.grid, rho0_minus.grid = (L, rho_plus), (L, rho_minus)
    return [rho0_plus, rho0_minus, rho_plus, rho_minus]


# ---- rank plumbing ---------------------------------------------------------------------------
def dist_info():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous block partition; the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_order(n_items: int, world: int) -> np.ndarray:
    """Replica order whose contiguous shard_bounds blocks are the strided sets {r, r+world, ...}: a sweep lists its
    replicas parameter-major and the event rate depends on the parameter (flip rate exp(-beta*sigma*m)), so contiguous
    blocks would give the ranks unequal work; strided blocks give every rank the same mix of sweep points."""
    return np.concatenate([np.arange(r, n_items, world) for r in range(world)]) if world > 1 else np.arange(n_items)


def replica_cost(spec: "EnsembleSpec") -> np.ndarray:
    """Relative cost estimate of every replica: its total event rate in the initial state,
    n (2D + lambda f_plus) + n_plus exp(-beta m0) + n_minus exp(+beta m0) with m0 the initial magnetisation (the rates
    of CLASS.py:276-319,59-60 without the exclusion factors).  Only the ORDER matters: it drives longest-first scheduling."""
    ps = spec.ps_kwargs
    L, K = int(ps["L"]), int(ps.get("site_capacity", 1))
    D, lam = float(ps["rate_diffusion"]), float(ps["rate_active"])
    if ps.get("scale_rates", True):
        dx = float(ps.get("xlim", 1.0)) / L
        D, lam = D / dx ** 2, lam / dx
    R = len(spec.betas)
    if ps.get("init", "fixed") == "poisson" and spec.profiles_plus is not None:
        rp, rm = np.asarray(spec.profiles_plus, float), np.asarray(spec.profiles_minus, float)
        tot = rp + rm
        occ = np.minimum(float(K), tot) if K > 1 else 1.0 - np.exp(-tot)            # expected particles per site (K = 1 exact)
        frac = np.divide(rp, tot, out=np.full_like(tot, 0.5), where=tot > 0)
        n_plus_p, n_minus_p = (occ * frac).sum(axis=1), (occ * (1.0 - frac)).sum(axis=1)
        which = np.asarray(spec.profile_of, int) if spec.profile_of is not None else np.zeros(R, int)
        n_plus, n_minus = n_plus_p[which], n_minus_p[which]
    else:
        n = np.asarray(spec.N_of, float) if spec.N_of is not None else np.full(R, float(ps.get("N", 1000)))
        n_plus = n_minus = 0.5 * n
    n = n_plus + n_minus
    m0 = np.divide(n_plus - n_minus, n, out=np.zeros(R), where=n > 0)
    b = np.asarray(spec.betas, float)
    return n * 2.0 * D + lam * n_plus + n_plus * np.exp(-b * m0) + n_minus * np.exp(b * m0)


def schedule_order(spec: "EnsembleSpec", world: int) -> np.ndarray:
    """Replica order for a launch over `world` ranks: strided rank assignment (balanced_order), then longest-first inside
    every rank's block — the replicas of a block fill the GPU in about two waves of CTAs, and starting the expensive
    ones first shortens the tail of the last wave."""
    n_items = len(spec.betas)
    order = balanced_order(n_items, world)
    cost = replica_cost(spec)
    for r in range(world):
        lo, hi = shard_bounds(n_items, r, world)
        blk = order[lo:hi]
        order[lo:hi] = blk[np.argsort(-cost[blk], kind="stable")]
    return order


def permute_spec(spec: "EnsembleSpec", order: np.ndarray) -> "EnsembleSpec":
    """The same ensemble with its per-replica arrays reordered (seeds travel with their replica, so results do not change)."""
    from dataclasses import replace
    pick = lambda a: None if a is None else np.asarray(a)[order]
    return replace(spec, betas=pick(spec.betas), point_of=pick(spec.point_of), seeds=pick(spec.seeds),
                   profile_of=pick(spec.profile_of), N_of=pick(spec.N_of))


def expected_poisson_particles(rho_p, rho_m, K):
    """mean and an upper bound of n for the K-truncated Poisson init (CLASS.py:160-189)."""
    lam = np.asarray(rho_p, dtype=float) + np.asarray(rho_m, dtype=float)
    pk = np.exp(-lam)                       # P(count = k), k = 0 .. K-1, vectorised over the sites
    cdf = np.zeros_like(lam)
    e = np.zeros_like(lam)
    for k in range(K):
        e += k * pk
        cdf += pk
        pk = pk * lam / (k + 1)
    mean = float((e + K * (1.0 - cdf)).sum())
    return mean, int(math.ceil(mean + 6.5 * math.sqrt(max(mean, 1.0)) + 8))


@dataclass
class EnsembleSpec:
    """One launch group: replicas that share the model parameters (L, K, D, lambda, sigma, T, obs_dt)."""
    ps_kwargs: dict
    run_kwargs: dict
    betas: np.ndarray                 # [R] beta of every replica (grid-point-major)
    point_of: np.ndarray              # [R] grid-point index of every replica
    seeds: np.ndarray                 # [R] uint64 Philox keys
    profiles_plus: np.ndarray | None = None    # [n_profiles][L] ('poisson')
    profiles_minus: np.ndarray | None = None
    profile_of: np.ndarray | None = None       # [R]
    N_of: np.ndarray | None = None             # [R] ('fixed' with per-point N)
    record: int = APS_REC_COUNTS | APS_REC_POS


@dataclass
class EnsembleResult:
    reducers: np.ndarray              # [R][APS_RED_N] (all ranks, original replica order)
    n_events: np.ndarray              # [R]
    status: np.ndarray                # [R]
    n_particles: np.ndarray           # [R]
    profiles: np.ndarray | None       # [points][4][L] sums over the point's replicas (all ranks)
    reps_per_point: np.ndarray | None
    m_hist: np.ndarray | None = None  # [points][256] histogram of the per-replica time-averaged magnetisation on [-1, 1] (all ranks)
    mbar: np.ndarray | None = None    # [R] the per-replica values themselves
    info: dict = field(default_factory=dict)


def _model_params(ps_kwargs):
    L = int(ps_kwargs["L"])
    xlim = float(ps_kwargs.get("xlim", 1.0))
    dx = xlim / L
    D, lam = float(ps_kwargs["rate_diffusion"]), float(ps_kwargs["rate_active"])
    if ps_kwargs.get("scale_rates", True):
        D, lam = D / dx ** 2, lam / dx
    sigma = float(ps_kwargs.get("local_kernel_sigma", 0.005))
    periodic = bool(ps_kwargs.get("periodic", False))
    if sigma > 0 and periodic:
        radius, weights = periodic_weights(L, dx, sigma)      # ring kernel, CLASS.py:111-121
    elif sigma > 0:
        radius, weights = gaussian_weights(sigma / dx)
    else:
        radius, weights = -1, np.zeros(1)
    if ps_kwargs.get("anchor_positions") is not None:
        raise NotImplementedError("anchor_positions: use ParticleSystem.run() (variable particle count); the ensemble "
                                  "launcher covers the anchor-free sweeps the drivers ship")
    from .engine import tabulate_flip_rate
    fn = ps_kwargs.get("flip_rate_fn")          # one callable for every replica of the sweep, as in the reference's loops
    return dict(flip_tab=None if fn is None else tabulate_flip_rate(fn), L=L, dx=dx, D=D, lam=lam, K=int(ps_kwargs.get("site_capacity", 1)), radius=radius, weights=weights,
                crowding=bool(ps_kwargs.get("crowding_suppresses_rates", False)), periodic=periodic, init=ps_kwargs.get("init", "fixed"),
                N=int(ps_kwargs.get("N", 1000)))


class DeviceEnsemble:
    """This rank's shard of an EnsembleSpec, resident in HBM."""

    def __init__(self, spec: EnsembleSpec, lo: int, hi: int, device=None):
        self.spec, self.lo, self.hi = spec, lo, hi
        self.mp = mp = _model_params(spec.ps_kwargs)
        self.dev = _dev(device)
        self.lib = capi.load()
        T, obs_dt = float(spec.run_kwargs.get("T", 10.0)), float(spec.run_kwargs.get("obs_dt", 0.01))
        self.times_obs = np.arange(0.0, T, obs_dt)
        self.T = T
        R = hi - lo
        self.R = R
        sl = slice(lo, hi)
        if mp["init"] == "poisson":
            self.n_max = max(expected_poisson_particles(spec.profiles_plus[p], spec.profiles_minus[p], mp["K"])[1]
                             for p in range(len(spec.profiles_plus)))
        else:
            self.n_max = int(spec.N_of.max()) if spec.N_of is not None else mp["N"]
        self.n_max = max(8, (self.n_max + 7) // 8 * 8)
        # ---- host -> device: everything the shard needs (this is the e2e H2D traffic) ----
        up = lambda a, dt: torch.as_tensor(np.array(a, dtype=dt, order="C", copy=True)).to(self.dev, non_blocking=True)
        self.h2d_bytes = 0

        def upc(a, dt):
            t = up(a, dt)
            self.h2d_bytes += t.numel() * t.element_size()
            return t

        self.seeds = upc(np.asarray(spec.seeds[sl], dtype=np.uint64).view(np.int64), np.int64)
        self.betas_h = np.asarray(spec.betas[sl], dtype=np.float64)
        self.rho_p = upc(spec.profiles_plus, np.float64) if mp["init"] == "poisson" else None
        self.rho_m = upc(spec.profiles_minus, np.float64) if mp["init"] == "poisson" else None
        self.profile_of = upc(spec.profile_of[sl], np.int32) if (mp["init"] == "poisson" and spec.profile_of is not None) else None
        self.N_of = upc(spec.N_of[sl], np.int32) if (mp["init"] == "fixed" and spec.N_of is not None) else None
        self.pos0 = torch.zeros((R, self.n_max), dtype=torch.int32, device=self.dev)
        self.sigma0 = torch.ones((R, self.n_max), dtype=torch.int8, device=self.dev)
        self.n = torch.zeros((R,), dtype=torch.int32, device=self.dev)
        self.rb = ReplicaBatch(L=mp["L"], K=mp["K"], radius=mp["radius"], weights=mp["weights"], D=mp["D"], lam=mp["lam"],
                               T=T, times_obs=self.times_obs, betas=self.betas_h, n=self.n, pos0=self.pos0,
                               sigma0=self.sigma0, seeds=self.seeds, record=spec.record, crowding=mp["crowding"],
                               device=self.dev.index, dx=mp["dx"], periodic=mp["periodic"], flip_tab=mp["flip_tab"],
                               zero_rows=bool(spec.record & capi.APS_REC_MLOCAL))   # reducer-only ensembles never read a row >= n_obs
        self.h2d_bytes += (self.rb.times_obs.numel() + self.rb.beta.numel() + (self.rb.weights.numel() if self.rb.weights is not None else 0)) * 8
        self.n_points = int(spec.point_of.max()) + 1 if len(spec.point_of) else 0
        # replica lists per grid point (CSR) for the device-side per-point sums: the shard is in schedule order
        pl = np.asarray(spec.point_of[sl], dtype=np.int64)
        by_point = np.argsort(pl, kind="stable")
        self.point_of = upc(pl, np.int32)
        self.point_reps = upc(by_point, np.int32)
        self.point_start = upc(np.searchsorted(pl[by_point], np.arange(self.n_points + 1)), np.int32)

    def init_particles(self):
        mp = self.mp
        a = ApsInitArgs(self.R, mp["L"], mp["K"], self.n_max, 1 if mp["init"] == "poisson" else 0, mp["N"],
                        int(self.rho_p.shape[0]) if self.rho_p is not None else 0, 0,
                        self.rho_p.data_ptr() if self.rho_p is not None else None,
                        self.rho_m.data_ptr() if self.rho_m is not None else None,
                        self.profile_of.data_ptr() if self.profile_of is not None else None,
                        self.N_of.data_ptr() if self.N_of is not None else None,
                        self.seeds.data_ptr(), self.pos0.data_ptr(), self.sigma0.data_ptr(), self.n.data_ptr())
        capi.check(self.lib.aps_init_particles_device(a, _stream()), "aps_init_particles_device")

    def step(self, want_profiles=True):
        """init -> K1 -> per-run reducers, magnetisation histogram (-> per-point profile sums).  Hand-written kernels
        only, all on the device, no sync."""
        self.init_particles()
        self.rb.run_philox()
        self.red = self.rb.reduce()
        self.hist, self.mbar = self.rb.m_histogram(max(1, self.n_points), self.point_of)     # [points][256] int64, [R]
        self.prof = None
        if want_profiles and self.n_points:
            self.prof = self.rb.profile_sums_by_point(self.n_points, self.point_start, self.point_reps)   # [points][4][L]
        return self

    def pack_scalars(self):
        """[R][APS_RED_N + 4] f64: reducers, n_events, status, n, time-averaged m — the only per-replica D2H payload."""
        rb = self.rb
        return torch.cat([self.red, rb.n_events.double()[:, None], rb.status.double()[:, None],
                          self.n.double()[:, None], self.mbar[:, None]], dim=1)


def run_ensemble(spec: EnsembleSpec, want_profiles=True, device=None, ensemble_cls=None) -> EnsembleResult:
    """Run every replica of `spec` across the ranks of the current process group and return the
    gathered result on every rank.  `ensemble_cls` lets the CPU tests substitute an oracle-backed
    shard for DeviceEnsemble to exercise the sharding / collective logic under gloo."""
    rank, world = dist_info()
    Rtot = len(spec.betas)
    lo, hi = shard_bounds(Rtot, rank, world)
    order = schedule_order(spec, world)            # rank r runs replicas r, r+world, ... (equal mix of sweep points), longest first
    ens = (ensemble_cls or DeviceEnsemble)(permute_spec(spec, order), lo, hi, device=device)
    ens.step(want_profiles=want_profiles)
    scal = ens.pack_scalars()
    prof = ens.prof
    hist = getattr(ens, "hist", None)
    if world > 1:
        width = scal.shape[1]
        most = shard_bounds(Rtot, 0, world)[1]
        pad = torch.zeros((most, width), dtype=torch.float64, device=scal.device)
        pad[: scal.shape[0]] = scal
        gathered = [torch.zeros_like(pad) for _ in range(world)]
        torch.distributed.all_gather(gathered, pad)
        parts = []
        for r in range(world):
            a, b = shard_bounds(Rtot, r, world)
            parts.append(gathered[r][: b - a])
        scal = torch.cat(parts, dim=0)
        if prof is not None:
            torch.distributed.all_reduce(prof, op=torch.distributed.ReduceOp.SUM)
        if hist is not None:
            torch.distributed.all_reduce(hist, op=torch.distributed.ReduceOp.SUM)
    scal_h = np.empty(tuple(scal.shape), dtype=np.float64)
    scal_h[order] = scal.cpu().numpy()             # back to the caller's replica order
    prof_h = prof.cpu().numpy() if prof is not None else None
    reps = np.bincount(np.asarray(spec.point_of, dtype=np.int64)) if len(spec.point_of) else None
    info = dict(rank=rank, world=world, shard=(lo, hi), n_max=ens.n_max, h2d_bytes=ens.h2d_bytes,
                d2h_bytes=int(scal.numel() * 8 + (prof.numel() * 8 if prof is not None else 0) +
                              (hist.numel() * 8 if hist is not None else 0)), M=ens.rb.M)
    n_part = scal_h[:, APS_RED_N + 2].astype(np.int64)
    if (n_part < 0).any():
        raise capi.ApsError("initial sample exceeded n_max; raise the particle bound")
    return EnsembleResult(reducers=scal_h[:, :APS_RED_N], n_events=scal_h[:, APS_RED_N].astype(np.int64),
                          status=scal_h[:, APS_RED_N + 1].astype(np.int32), n_particles=n_part, profiles=prof_h,
                          reps_per_point=reps, info=info, m_hist=hist.cpu().numpy() if hist is not None else None,
                          mbar=scal_h[:, APS_RED_N + 3] if scal_h.shape[1] > APS_RED_N + 3 else None)


# ---- reference-shaped entry points -------------------------------------------------------------
def _profiles_from_init_kwargs(ps_kwargs, init_kwargs):
    L = int(ps_kwargs["L"])
    def tabulate(fn):                                                                    # CLASS.py:71-72: fn(i / L), i = 0..L-1
        grid = getattr(fn, "grid", None)
        if grid is not None and grid[0] == L:      # a make_exp_gradient callable: same values, evaluated vectorised
            idx = np.clip(np.round((np.arange(L) / L) * L), 0, L - 1).astype(int)
            return np.asarray(grid[1], dtype=float)[idx]
        return np.array([fn(i / L) for i in range(L)], dtype=float)
    return tabulate(init_kwargs["rho0_plus"]), tabulate(init_kwargs["rho0_minus"])


def _mean_std_se(x):
    x = np.asarray(x, dtype=float)
    mean = float(x.mean())
    std = float(x.std(ddof=1)) if x.size > 1 else 0.0
    return mean, std, std / np.sqrt(max(1, x.size))


def replica_seeds(point_of, run_idx, base_seed=None):
    """Philox keys of the replicas (they also key the device-side initial conditions).
    base_seed=None (the default of every sweep entry point): fresh OS entropy, independent streams for every replica
    and every call — what the reference does with `rng=None` (sweep_beta.py:77-83 -> default_rng(), CLASS.py:75-76).
    base_seed=int: reproducible, seed = base + stride * point + run with stride = max(10 000, runs per point), so
    that keys cannot collide inside a sweep (seed = 10 000 * beta_idx + replica_idx for config 2)."""
    point_of, run_idx = np.asarray(point_of, dtype=np.uint64), np.asarray(run_idx, dtype=np.uint64)
    if base_seed is None:
        return np.random.SeedSequence().generate_state(len(point_of), np.uint64)
    stride = np.uint64(max(10_000, int(run_idx.max()) + 1 if len(run_idx) else 1))
    return np.uint64(base_seed) + stride * point_of + run_idx


def build_beta_sweep_spec(beta_values, n_runs_per_beta, ps_kwargs, init_kwargs, run_kwargs, base_seed=None):
    betas = np.repeat(np.asarray(beta_values, dtype=float), n_runs_per_beta)
    point_of = np.repeat(np.arange(len(beta_values)), n_runs_per_beta)
    run_idx = np.tile(np.arange(n_runs_per_beta), len(beta_values))
    seeds = replica_seeds(point_of, run_idx, base_seed)
    ps = dict(ps_kwargs)
    kw = {}
    if ps.get("init", "fixed") == "poisson":
        rp, rm = _profiles_from_init_kwargs(ps, init_kwargs)
        kw = dict(profiles_plus=rp[None], profiles_minus=rm[None])
    return EnsembleSpec(ps_kwargs=ps, run_kwargs=dict(run_kwargs), betas=betas, point_of=point_of, seeds=seeds, **kw)


def sweep_over_betas(beta_values, n_runs_per_beta=10, ps_kwargs=None, init_kwargs=None, run_kwargs=None,
                     base_seed=None, want_profiles=True, ensemble_cls=None):
    """All (beta, run) replicas in one sharded launch; returns the arrays the reference's
    sweep_over_betas collects (sweep_beta.py:880-931), keyed like its save_dict."""
    spec = build_beta_sweep_spec(beta_values, n_runs_per_beta, ps_kwargs or {}, init_kwargs or {}, run_kwargs or {},
                                 base_seed=base_seed)
    res = run_ensemble(spec, want_profiles=want_profiles, ensemble_cls=ensemble_cls)
    nb = len(beta_values)
    red = res.reducers.reshape(nb, n_runs_per_beta, APS_RED_N)
    out = dict(beta_values=np.asarray(beta_values, dtype=float), raw_by_beta=[red[b, :, capi.APS_RED_V_EFF].copy() for b in range(nb)])
    # key names of the reference's pre_dict / save_dict (sweep_beta.py:952-970,1002-1026)
    for stem, col in [("", capi.APS_RED_V_EFF), ("D_", capi.APS_RED_D_EFF), ("m_", capi.APS_RED_M_MEAN),
                      ("rho_", capi.APS_RED_RHO_EFF), ("block_", capi.APS_RED_BLOCK)]:
        x = red[:, :, col]                                   # [beta][run]; same statistics as _mean_std_se per row
        out[stem + "means"] = x.mean(axis=1)
        out[stem + "stds"] = x.std(axis=1, ddof=1) if n_runs_per_beta > 1 else np.zeros(nb)
        out[stem + "ses"] = out[stem + "stds"] / np.sqrt(max(1, n_runs_per_beta))
    out["ps_kwargs"] = dict(ps_kwargs or {})
    out["outs"] = []          # per-run dicts stay on the device; the reducers above replace them
    out["n_events"] = res.n_events.reshape(nb, n_runs_per_beta)
    out["status"] = res.status.reshape(nb, n_runs_per_beta)
    if res.m_hist is not None:      # device-side histogram of the per-run time-averaged magnetisation, 256 bins on [-1, 1]
        out["m_hist"], out["m_hist_edges"] = res.m_hist, np.linspace(-1.0, 1.0, res.m_hist.shape[1] + 1)
    if res.mbar is not None:
        out["m_bar"] = res.mbar.reshape(nb, n_runs_per_beta)
    if res.profiles is not None:
        reps = float(n_runs_per_beta)
        out["rho_plus_profile_mean"] = res.profiles[:, 0] / reps
        out["rho_minus_profile_mean"] = res.profiles[:, 1] / reps
        var_p = np.maximum(res.profiles[:, 2] / reps - (res.profiles[:, 0] / reps) ** 2, 0.0)
        out["rho_plus_profile_se"] = np.sqrt(var_p / max(1.0, reps - 1.0))
    out["info"] = res.info
    return out


def sweep_beta_ensemble(beta, n_runs=10, ps_kwargs=None, init_kwargs=None, run_kwargs=None, rng_seeds=None):
    """Drop-in for sweep_beta.py:56-117 (same positional return tuple; `out_list` is empty because
    the per-run reducers already ran on the device)."""
    spec = build_beta_sweep_spec([beta], n_runs, ps_kwargs or {}, init_kwargs or {}, run_kwargs or {})
    if rng_seeds is not None:
        spec.seeds = np.asarray([int(s) for s in rng_seeds[:n_runs]], dtype=np.uint64)
    res = run_ensemble(spec, want_profiles=False)
    r = res.reducers
    mean, std, se = _mean_std_se(r[:, capi.APS_RED_V_EFF])
    D = _mean_std_se(r[:, capi.APS_RED_D_EFF]); m = _mean_std_se(r[:, capi.APS_RED_M_MEAN])
    rho = _mean_std_se(r[:, capi.APS_RED_RHO_EFF]); blk = _mean_std_se(r[:, capi.APS_RED_BLOCK])
    return (mean, std, se, r[:, capi.APS_RED_V_EFF].copy(), [], m[0], m[1], m[2], rho[0], rho[2], blk[0], blk[2], D[0], D[2])


def init_distributed_from_env():
    """One process per GPU, launched by torchrun: bind LOCAL_RANK's device and join the NCCL group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        capi.check(capi.load().aps_set_device(local), "aps_set_device")
    if world > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo")
    return int(os.environ.get("RANK", "0")), world


# ---- (N_part, beta) double sweep: ..._double_sweep.py:851-873 -------------------------------------
def build_double_sweep_spec(n_part_values, beta_values, n_runs, ps_kwargs, run_kwargs, frac_plus=0.75, decay_plus=0.2,
                            decay_minus=0.2, base_seed=None):
    """Grid points ordered N_part-major, beta-minor (the order of the reference's nested loops); every
    density gets its own Poisson intensity profile (make_exp_gradient with N = N_part)."""
    n_part_values = [int(v) for v in n_part_values]
    nb = len(beta_values)
    L = int(ps_kwargs["L"])
    prof_p = np.stack([make_exp_gradient(L=L, N=N, frac_plus=frac_plus, decay_length=decay_plus, anchor_positions=None)[2]
                       for N in n_part_values])
    prof_m = np.stack([make_exp_gradient(L=L, N=N, frac_plus=frac_plus, decay_length=decay_minus, anchor_positions=None)[3]
                       for N in n_part_values])
    point = np.arange(len(n_part_values) * nb)
    betas = np.repeat(np.tile(np.asarray(beta_values, dtype=float), len(n_part_values)), n_runs)
    point_of = np.repeat(point, n_runs)
    profile_of = np.repeat(np.repeat(np.arange(len(n_part_values)), nb), n_runs).astype(np.int32)
    run_idx = np.tile(np.arange(n_runs), len(point))
    seeds = replica_seeds(point_of, run_idx, base_seed)
    ps = dict(ps_kwargs, init="poisson")
    return EnsembleSpec(ps_kwargs=ps, run_kwargs=dict(run_kwargs), betas=betas, point_of=point_of, seeds=seeds,
                        profiles_plus=prof_p, profiles_minus=prof_m, profile_of=profile_of)


def double_sweep(n_part_values, beta_values, n_runs, ps_kwargs, run_kwargs, **kw):
    """Returns {N_part: dict like sweep_over_betas' save_dict} for every density of the grid."""
    spec = build_double_sweep_spec(n_part_values, beta_values, n_runs, ps_kwargs, run_kwargs, **kw)
    res = run_ensemble(spec, want_profiles=False)
    nb = len(beta_values)
    red = res.reducers.reshape(len(n_part_values), nb, n_runs, APS_RED_N)
    out = {}
    for i, N in enumerate(n_part_values):
        d = dict(beta_values=np.asarray(beta_values, dtype=float))
        for stem, col in [("", capi.APS_RED_V_EFF), ("D_", capi.APS_RED_D_EFF), ("m_", capi.APS_RED_M_MEAN),
                          ("rho_", capi.APS_RED_RHO_EFF), ("block_", capi.APS_RED_BLOCK)]:
            stats = [_mean_std_se(red[i, b, :, col]) for b in range(nb)]
            d[stem + "means"] = np.array([s[0] for s in stats])
            d[stem + "stds"] = np.array([s[1] for s in stats])
            d[stem + "ses"] = np.array([s[2] for s in stats])
        out[int(N)] = d
    out["info"] = dict(res.info, n_events_total=int(res.n_events.sum()))
    return out


# ---- structure observables: PARTICLE_solver_BIOLOGY_local_structure.py:55-193 -----------------------
from .structure import structure_observables, fft_amplitudes      # noqa: E402  (device-side analyses live in structure.py)


def sweep_betas_for_structures(beta_values, n_runs_per_beta, ps_kwargs, init_kwargs, run_kwargs, start_fraction=0.5,
                               k_max=None, base_seed=None, keep_raw=True, k_keep=64):
    """Drop-in for local_structure.py:167-193: dict keyed by beta with the ensemble keys of
    sweep_beta_structure_ensemble (:105-165).  `raw` holds one entry per run with the per-run observables and a light
    `out` dict (`times_obs`, `fft_amp_list` restricted to the first `k_keep` modes, `var_list`, `m_global`) — enough
    for the driver's time-series analyses (time_to_pattern, lowk_variance_time, extract_growth_rate; k <= 25 there)
    without downloading the (M, L) arrays.
    Multi-GPU: the beta values are dealt to the ranks in strides (rank r takes betas r, r+world, ...: the event rate grows
    with beta), every rank runs its betas in one launch and analyses them on its device; the per-beta result dicts (a few
    KB each) are exchanged with one all_gather_object, so every rank returns the full dict."""
    from .capi import APS_REC_MLOCAL
    rank, world = dist_info()
    beta_values = list(beta_values)
    seeds_all = replica_seeds(np.repeat(np.arange(len(beta_values)), n_runs_per_beta),
                              np.tile(np.arange(n_runs_per_beta), len(beta_values)), base_seed)
    if world > 1 and base_seed is None:          # fresh entropy must be the same on every rank: rank 0 decides
        box = [seeds_all]
        torch.distributed.broadcast_object_list(box, src=0)
        seeds_all = box[0]
    mine = list(range(rank, len(beta_values), world))
    results_local = {}
    if mine:
        spec = build_beta_sweep_spec([beta_values[i] for i in mine], n_runs_per_beta, ps_kwargs, init_kwargs, run_kwargs, base_seed=0)
        spec.seeds = np.concatenate([seeds_all[i * n_runs_per_beta:(i + 1) * n_runs_per_beta] for i in mine])
        spec.record = APS_REC_COUNTS | APS_REC_POS | APS_REC_MLOCAL
        ens = DeviceEnsemble(spec, 0, len(spec.betas))
        ens.init_particles()
        ens.rb.run_philox()
        amp, total, var = fft_amplitudes(ens.rb)
        obs = {k: v.cpu().numpy() for k, v in structure_observables(ens.rb, start_fraction, k_max, amp=amp, var=var).items()}
        nr = n_runs_per_beta
        n_events = ens.rb.n_events.cpu().numpy()
        if keep_raw:
            amp_h = amp[:, :, :k_keep].cpu().numpy()
            var_h = var.cpu().numpy()
            n_h = np.maximum(1, ens.n.cpu().numpy()).astype(float)
            mg_h = ens.rb.obs_sigma_sum.cpu().numpy() / n_h[:, None]
        for b, i in enumerate(mine):
            sl = slice(b * nr, (b + 1) * nr)
            se = lambda x: x.std(ddof=1) / np.sqrt(nr)
            res = {
                "var_mean": obs["var_mean"][sl].mean(), "var_se": se(obs["var_mean"][sl]),
                "low_k_power_mean": obs["low_k_power"][sl].mean(), "low_k_power_se": se(obs["low_k_power"][sl]),
                "dominant_k_mode": int(np.round(obs["dominant_k"][sl].mean())),
                "m_local_var_mean": obs["m_local_var"][sl].mean(), "m_local_var_se": se(obs["m_local_var"][sl]),
                "fft_mean_mean": obs["fft_mean"][sl].mean(axis=0),
                "fft_mean_se": obs["fft_mean"][sl].std(axis=0, ddof=1) / np.sqrt(nr),
                "lowk_var_mean": obs["lowk_variance"][sl].mean(), "lowk_var_se": se(obs["lowk_variance"][sl]),
                "n_events": n_events[sl],
            }
            if keep_raw:
                res["raw"] = [
                    dict({k: (obs[k][r] if obs[k].ndim > 1 else obs[k][r].item()) for k in obs},
                         out=dict(times_obs=ens.times_obs, fft_amp_list=amp_h[r], var_list=var_h[r], m_global=mg_h[r]))
                    for r in range(b * nr, (b + 1) * nr)]
            results_local[i] = res
    if world > 1:
        parts = [None] * world
        torch.distributed.all_gather_object(parts, results_local)
        results_local = {i: r for part in parts for i, r in part.items()}
    return {beta_values[i]: results_local[i] for i in range(len(beta_values))}


def sweep_over_sigmas(sigma_values, beta_values, n_runs_per_beta=5, ps_kwargs=None, init_kwargs=None, run_kwargs=None,
                      base_seed=None, save_dir=None):
    """Drop-in for sweep_beta_2.py:1030-1075: one beta sweep per interaction width sigma (each sigma is its own launch
    group: the filter radius differs), results keyed by sigma with the reference's keys; with `save_dir` every sigma
    is also written to `v_eff_vs_beta_sigma_<sigma>.npz` under the reference's key names (:1059-1067)."""
    import os
    results = {}
    for si, sigma in enumerate(sigma_values):
        ps = dict(ps_kwargs or {}, local_kernel_sigma=float(sigma))
        # every sigma gets its own streams (a reproducible base_seed is offset per sigma; None = fresh entropy per sweep)
        seed_s = None if base_seed is None else int(base_seed) + si * 1_000_003 * max(10_000, n_runs_per_beta)
        sd = sweep_over_betas(beta_values, n_runs_per_beta, ps, init_kwargs, run_kwargs, base_seed=seed_s,
                              want_profiles=False)
        results[sigma] = {"beta": np.asarray(beta_values, dtype=float), "v_mean": sd["means"], "v_se": sd["ses"],
                          "D_mean": sd["D_means"], "D_se": sd["D_ses"], "ps_kwargs": sd["ps_kwargs"]}
        if save_dir is not None:
            keep = {k: v for k, v in sd["ps_kwargs"].items() if not callable(v)}
            np.savez(os.path.join(save_dir, f"v_eff_vs_beta_sigma_{sigma:.4g}.npz"), beta=results[sigma]["beta"],
                     v_mean=sd["means"], v_se=sd["ses"], D_mean=sd["D_means"], D_se=sd["D_ses"],
                     ps_kwargs=np.array(keep, dtype=object))
    return results


# ---- result caching with the reference's npz key names (sweep_beta.py:933-970) ----------------------
NPZ_KEYS = ["beta_values", "means", "stds", "ses", "D_means", "D_ses", "m_means", "m_stds", "m_ses", "rho_means",
            "rho_ses", "block_means", "block_ses", "ps_kwargs", "outs"]


def save_sweep_npz(path, sweep_out):
    """np.savez of a sweep_over_betas result under the key names the reference's `run=False` reload path reads
    (`save_dict['means']`, ..., `save_dict['ps_kwargs'].item()`, sweep_beta.py:933-950)."""
    ps = {k: v for k, v in dict(sweep_out.get("ps_kwargs", {})).items() if not callable(v)}
    payload = {k: sweep_out[k] for k in NPZ_KEYS if k in sweep_out and k not in ("ps_kwargs", "outs")}
    payload["ps_kwargs"] = np.array(ps, dtype=object)
    payload["outs"] = np.array(sweep_out.get("outs", []), dtype=object)
    np.savez(path, **payload)
    return path


def load_sweep_npz(path):
    """The reference's reload (`data = np.load(..., allow_pickle=True); save_dict = dict(data)`)."""
    data = np.load(path, allow_pickle=True)
    d = dict(data)
    d["ps_kwargs"] = d["ps_kwargs"].item()
    return d
