"""Device-side plumbing shared by the drop-in class and the ensemble launcher.

`ReplicaBatch` owns the HBM buffers of one batch of independent replicas (torch tensors used
purely as device memory) and drives the C ABI (`include/aps.h`) on the current CUDA stream.
HBM layout, per batch of R replicas (L sites, M observation rows, n_max particle slots):
    pos0/pos_end  int32 [R][n_max]     sigma0/sigma_end int8 [R][n_max]
    obs_cp/obs_cm int8  [R][M][L]      obs_pos int32 [R][M][n_max]    obs_sigma_sum int32 [R][M]
    obs_m_local   f64   [R][M][L] (only when asked for)
    per-replica scalars: n, beta, seeds, n_obs, n_events, t_end, status, n_guard, draws_used
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi
from .batch import make_batch, make_params
from .capi import (APS_REC_COUNTS, APS_REC_MLOCAL, APS_REC_POS, APS_RED_N, ApsExpandArgs, ApsHistArgs, ApsProfileArgs,
                   ApsReduceArgs)


def gaussian_weights(sigma_grid: float):
    """Taps of scipy.ndimage.gaussian_filter1d(sigma=sigma_grid, truncate=4.0) exactly as scipy builds
    them (_gaussian_kernel1d, order 0), which is what compute_local_m_field uses (CLASS.py:229-238)."""
    sd = float(sigma_grid)
    radius = int(4.0 * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return radius, phi[::-1].copy()


FLIP_TABLE_G = 8192


def tabulate_flip_rate(flip_rate_fn, G=FLIP_TABLE_G):
    """A custom `flip_rate_fn(sigma, m)` (CLASS.py:59-62; called as fn(sigma[n] int8, m[n] float64) at :262) evaluated on
    the host on the grid m_k = -1 + 2k/G for sigma = +1 and sigma = -1 -> float64 [2][G+1], the table the kernels
    interpolate linearly (include/aps_math.h aps_flip_interp).  The callable never runs on the device."""
    grid = -1.0 + 2.0 * np.arange(G + 1, dtype=np.float64) / G
    rows = [np.asarray(flip_rate_fn(np.full(G + 1, sg, dtype=np.int8), grid), dtype=np.float64) for sg in (1, -1)]
    tab = np.stack([np.broadcast_to(r, (G + 1,)) for r in rows]).copy()
    if not np.isfinite(tab).all() or (tab < 0).any():
        raise ValueError("flip_rate_fn must return finite, non-negative rates on m in [-1, 1]")
    return tab


def periodic_weights(L: int, dx: float, sigma: float, tail: float = 1e-22):
    """Taps of the reference's ring kernel (CLASS.py:111-121: exp(-0.5*(min(j, L-j)*dx/sigma)^2), normalised over the
    ring), truncated at the smallest radius whose discarded mass is below `tail`.  The reference applies the kernel
    by FFT (:224-227); the direct sum over the kept taps differs from it only by rounding (~1e-16) plus `tail`."""
    j = np.arange(L)
    kernel = np.exp(-0.5 * (np.minimum(j, L - j) * dx / sigma) ** 2)
    kernel = kernel / kernel.sum()
    half = (L - 1) // 2
    for r in range(0, half + 1):
        if kernel[r + 1:L - r].sum() <= tail:
            return r, np.concatenate([kernel[r:0:-1], kernel[:r + 1]]).copy()
    raise NotImplementedError("periodic=True: the kernel does not decay within half the ring "
                              "(local_kernel_sigma too large for this L)")


def _dev(device=None):
    if not torch.cuda.is_available():
        raise capi.ApsError("no CUDA device: the B200 stepper has no CPU path")
    return torch.device("cuda", torch.cuda.current_device() if device is None else device)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class ReplicaBatch:
    def __init__(self, *, L, K, radius, weights, D, lam, T, times_obs, betas, n, pos0, sigma0, seeds=None,
                 record=APS_REC_COUNTS | APS_REC_POS, crowding=False, device=None, dx=None, anchor_mask=None,
                 k_on=0.0, k_off=0.0, k_exit=0.0, suppress_flip_when_bound=True, immobilize_when_anchored=True,
                 exit_cap=None, periodic=False, flip_tab=None, zero_rows=True):
        self.lib = capi.load()
        self.dev = _dev(device)
        self.L, self.K, self.radius = int(L), int(K), int(radius)
        self.dx = float(dx) if dx is not None else 1.0 / self.L
        self.T = float(T)
        flags = capi.APS_FLAG_CROWDING if crowding else 0
        flags |= capi.APS_FLAG_PERIODIC if periodic else 0
        if anchor_mask is not None:
            flags |= (capi.APS_FLAG_SUPPRESS_FLIP_BOUND if suppress_flip_when_bound else 0)
            flags |= (capi.APS_FLAG_IMMOBILIZE if immobilize_when_anchored else 0)
        self.params = make_params(L, K, radius, D, lam, T, flags, k_on=k_on, k_off=k_off, k_exit=k_exit)
        t = lambda a, dt: torch.as_tensor(np.array(a, dtype=dt, order="C", copy=True)).to(self.dev, non_blocking=True)
        self.times_obs = t(times_obs, np.float64)
        self.M = int(self.times_obs.numel())
        self.weights = t(weights, np.float64) if radius >= 0 else None
        self.weights_host = np.array(weights, dtype=np.float64, order="C", copy=True) if radius >= 0 else None   # constant-bank taps of K1
        self.beta = t(betas, np.float64)
        self.R = int(self.beta.numel())
        self.n = n if isinstance(n, torch.Tensor) else t(n, np.int32)
        self.pos0 = pos0 if isinstance(pos0, torch.Tensor) else t(pos0, np.int32)
        self.sigma0 = sigma0 if isinstance(sigma0, torch.Tensor) else t(sigma0, np.int8)
        self.pos0 = self.pos0.reshape(self.R, -1).contiguous()
        self.sigma0 = self.sigma0.reshape(self.R, -1).contiguous()
        self.n_max = int(self.pos0.shape[1])
        self.seeds = None if seeds is None else (seeds if isinstance(seeds, torch.Tensor) else
                                                 t(np.asarray(seeds, dtype=np.uint64).view(np.int64), np.int64))
        self.record = int(record)
        # custom flip_rate_fn tabulated by the caller on m_k = -1 + 2k/G (see tabulate_flip_rate): [2][G+1]
        self.flip_tab = None if flip_tab is None else t(np.asarray(flip_tab, dtype=np.float64).reshape(2, -1), np.float64)
        self.flip_G = 0 if flip_tab is None else int(self.flip_tab.shape[1]) - 1
        R, M, Lq, nm, dv = self.R, self.M, self.L, self.n_max, self.dev
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dv)
        # observation rows: zero-filled like the reference's preallocated arrays (rows never reached stay zero, CLASS.py:465-472) unless
        # the caller only feeds them to the device-side reducers, which never read a row >= n_obs (saves a multi-GB memset per sweep)
        zr = z if zero_rows else (lambda shape, dt: torch.empty(shape, dtype=dt, device=dv))
        self.obs_cp = zr((R, M, Lq), torch.int8) if record & APS_REC_COUNTS else None
        self.obs_cm = zr((R, M, Lq), torch.int8) if record & APS_REC_COUNTS else None
        self.obs_pos = zr((R, M, nm), torch.int32) if record & APS_REC_POS else None
        self.obs_m_local = zr((R, M, Lq), torch.float64) if record & APS_REC_MLOCAL else None
        self.obs_sigma_sum = z((R, M), torch.int32)
        self.n_obs = z((R,), torch.int32)
        self.n_events = z((R,), torch.int64)
        self.t_end = z((R,), torch.float64)
        self.status = z((R,), torch.int32)
        self.n_guard = z((R,), torch.int64)
        self.draws_used = z((R,), torch.int64)
        self.pos_end = z((R, nm), torch.int32)
        self.sigma_end = z((R, nm), torch.int8)
        self.n_end = z((R,), torch.int32)
        self.obs_n = z((R, M), torch.int32)
        # anchors / binding / exit (only allocated when the batch has anchors)
        self.anchor_mask = None
        self.bound0 = self.bound_end = self.obs_bound = self.exit_t = self.exit_pos = self.n_exit = None
        self.exit_cap = 0
        if anchor_mask is not None:
            self.anchor_mask = t(np.asarray(anchor_mask, dtype=np.uint8), np.uint8)
            self.bound_end = z((R, nm), torch.int8)
            self.obs_bound = z((R, M, nm), torch.int8)
            self.exit_cap = int(exit_cap if exit_cap is not None else nm)
            self.exit_t = z((R, max(1, self.exit_cap)), torch.float64)
            self.exit_pos = z((R, max(1, self.exit_cap)), torch.int32)
            self.n_exit = z((R,), torch.int32)

    # -- K1 ---------------------------------------------------------------------------------
    def _batch(self, **extra):
        return make_batch(
            self.R, self.n_max, self.M, record=self.record, max_events=extra.pop("max_events", 0),
            spec_from=extra.pop("spec_from", -1), exit_cap=self.exit_cap, n_end=self.n_end, obs_n=self.obs_n,
            anchor_mask=self.anchor_mask, bound0=self.bound0, bound_end=self.bound_end, obs_bound=self.obs_bound,
            exit_t=self.exit_t, exit_pos=self.exit_pos, n_exit=self.n_exit, flip_tab=self.flip_tab, flip_G=self.flip_G,
            weights_host=self.weights_host,
            times_obs=self.times_obs, weights=self.weights, beta=self.beta, n=self.n, pos0=self.pos0,
            sigma0=self.sigma0, obs_cp=self.obs_cp, obs_cm=self.obs_cm, obs_pos=self.obs_pos,
            obs_sigma_sum=self.obs_sigma_sum, obs_m_local=self.obs_m_local, n_obs=self.n_obs,
            n_events=self.n_events, t_end=self.t_end, status=self.status, n_guard=self.n_guard,
            draws_used=self.draws_used, pos_end=self.pos_end, sigma_end=self.sigma_end, **extra)

    def run_philox(self, max_events=0, resume=None):
        """Native mode: in-kernel Philox4x32-10 streams keyed by `seeds`."""
        if self.seeds is None:
            raise ValueError("native mode needs seeds")
        extra = dict(seeds=self.seeds, max_events=max_events)
        if resume:
            extra.update(resume)
        b, keep = self._batch(**extra)
        capi.check(self.lib.aps_run_philox_device(self.params, b, _stream()), "aps_run_philox_device")
        return self

    def run_replay(self, draws, draw_off, max_events=0, resume=None, spec_from=-1):
        """Replay mode: consume the injected variate log (device tensors)."""
        extra = dict(draws=draws, draw_off=draw_off, max_events=max_events, spec_from=spec_from)
        if resume:
            extra.update(resume)
        b, keep = self._batch(**extra)
        capi.check(self.lib.aps_run_replay_device(self.params, b, _stream()), "aps_run_replay_device")
        return self

    # -- K4 ---------------------------------------------------------------------------------
    def expand(self, want_var=False):
        """rho_plus, rho_minus, total [R][M][L] (+ var [R][M]) on the device."""
        R, M, L = self.R, self.M, self.L
        z = lambda shape: torch.zeros(shape, dtype=torch.float64, device=self.dev)
        rho_p, rho_m, total = z((R, M, L)), z((R, M, L)), z((R, M, L))
        var = z((R, M)) if want_var else None
        a = ApsExpandArgs(R, M, L, 0, self.dx, self.n.data_ptr(), self.n_obs.data_ptr(), self.obs_cp.data_ptr(),
                          self.obs_cm.data_ptr(), rho_p.data_ptr(), rho_m.data_ptr(), total.data_ptr(),
                          var.data_ptr() if want_var else None, self.obs_n.data_ptr())
        capi.check(self.lib.aps_expand_obs_device(a, _stream()), "aps_expand_obs_device")
        return rho_p, rho_m, total, var

    def reduce(self, boundary_xmin=0.99, max_boundary_fraction=0.06, min_window_fraction=0.10,
               window_fraction=0.05, want_v_eff=False):
        """Per-run reducers -> tensor [R][APS_RED_N] (see include/aps.h APS_RED_*)."""
        out = torch.empty((self.R, APS_RED_N), dtype=torch.float64, device=self.dev)      # every slot is written by the kernel
        v = torch.zeros((self.R, self.M), dtype=torch.float64, device=self.dev) if want_v_eff else None
        a = ApsReduceArgs(self.R, self.M, self.L, self.n_max, self.dx, boundary_xmin, max_boundary_fraction,
                          min_window_fraction, window_fraction, self.times_obs.data_ptr(), self.n.data_ptr(),
                          self.n_obs.data_ptr(), self.obs_cp.data_ptr(), self.obs_cm.data_ptr(),
                          self.obs_pos.data_ptr() if self.obs_pos is not None else None,
                          self.obs_sigma_sum.data_ptr(), out.data_ptr(), v.data_ptr() if want_v_eff else None)
        capi.check(self.lib.aps_reduce_runs_device(a, _stream()), "aps_reduce_runs_device")
        return (out, v) if want_v_eff else out

    def profile_sums(self, reps_per_point, row_lo=None, row_hi=None):
        """[n_points][4][L]: sums over a point's replicas of the time-averaged rho_+, rho_- and their squares."""
        assert self.R % reps_per_point == 0
        P = self.R // reps_per_point
        row_lo = self.M // 2 if row_lo is None else row_lo
        row_hi = self.M if row_hi is None else row_hi
        prof = torch.zeros((P, 4, self.L), dtype=torch.float64, device=self.dev)
        a = ApsProfileArgs(P, reps_per_point, self.M, self.L, row_lo, row_hi, self.dx, self.n.data_ptr(),
                           self.n_obs.data_ptr(), self.obs_cp.data_ptr(), self.obs_cm.data_ptr(), prof.data_ptr())
        capi.check(self.lib.aps_profile_sums_device(a, _stream()), "aps_profile_sums_device")
        return prof

    def profile_sums_by_point(self, n_points, point_start, point_reps, row_lo=None, row_hi=None):
        """Same sums with an explicit replica list per grid point (int32 device tensors, CSR layout): the replicas of a
        rank's shard come in schedule order, not grid-point-major."""
        row_lo = self.M // 2 if row_lo is None else row_lo
        row_hi = self.M if row_hi is None else row_hi
        prof = torch.empty((n_points, 4, self.L), dtype=torch.float64, device=self.dev)
        if getattr(self, "_prof_scratch", None) is None:
            self._prof_scratch = torch.empty((self.R, 4, self.L), dtype=torch.float64, device=self.dev)
        a = ApsProfileArgs(n_points, 0, self.M, self.L, row_lo, row_hi, self.dx, self.n.data_ptr(), self.n_obs.data_ptr(),
                           self.obs_cp.data_ptr(), self.obs_cm.data_ptr(), prof.data_ptr(), point_start.data_ptr(),
                           point_reps.data_ptr(), self._prof_scratch.data_ptr(), self.R, 0)
        capi.check(self.lib.aps_profile_sums_device(a, _stream()), "aps_profile_sums_device")
        return prof

    def m_histogram(self, n_points=1, point_of=None, n_bins=256, lo=-1.0, hi=1.0, row_lo=None, row_hi=None):
        """Per-grid-point histogram (int64 [n_points][n_bins]) of the time-averaged magnetisation of every replica and
        the per-replica values (f64 [R]); device-side accumulation (`aps_m_histogram_device`)."""
        row_lo = self.M // 2 if row_lo is None else row_lo
        row_hi = self.M if row_hi is None else row_hi
        hist = torch.empty((n_points, n_bins), dtype=torch.int64, device=self.dev)       # zeroed by the call (accumulate = 0)
        mbar = torch.empty((self.R,), dtype=torch.float64, device=self.dev)
        obs_n = getattr(self, "obs_n", None)
        a = ApsHistArgs(self.R, self.M, n_points, n_bins, row_lo, row_hi, 0, 0, lo, hi, self.n.data_ptr(), self.n_obs.data_ptr(),
                        self.obs_sigma_sum.data_ptr(), obs_n.data_ptr() if obs_n is not None else None,
                        point_of.data_ptr() if point_of is not None else None, mbar.data_ptr(), hist.data_ptr())
        capi.check(self.lib.aps_m_histogram_device(a, _stream()), "aps_m_histogram_device")
        return hist, mbar
