"""Per-run structure analyses of PARTICLE_solver_BIOLOGY_local_structure.py, batched over the replicas of a
ReplicaBatch and evaluated on the device, so that the (M, L) observation arrays never travel to the host.

Reference functions mirrored (argument names and meaning kept; every function takes a leading replica axis):
  extract_structure_observables_from_out  local_structure.py:55-103   -> structure_observables
  time_to_pattern / ensemble_time_to_pattern            :195-209      -> time_to_pattern / ensemble_time_to_pattern
  cluster_size_distribution                              :210-222      -> cluster_size_distribution (host, one profile)
  temporal_autocorrelation                               :223-231      -> temporal_autocorrelation
  lowk_variance_time                                     :232-234      -> lowk_variance_time
  spectral_entropy / mode_competition_ratio              :235-245      -> same names (torch or numpy input)
  extract_growth_rate                                    :246-265      -> extract_growth_rate

|FFT| comes from cuFFT (torch.fft) on the density rows produced by the K4 expansion kernel; results agree with the
numpy versions to ~1e-12 relative (different FFT and summation order), which is the tolerance the tests state.
"""
from __future__ import annotations

import numpy as np
import torch


def fft_amplitudes(rb, k_keep=None):
    """|fft(total_list)| of every replica (CLASS.py:527-535): returns (amp [R][M][k], total [R][M][L], var [R][M])."""
    _, _, total, var = rb.expand(want_var=True)
    amp = torch.fft.fft(total, dim=-1).abs()
    if k_keep is not None:
        amp = amp[:, :, :k_keep]
    return amp, total, var


def structure_observables(rb, start_fraction=0.5, k_max=None, amp=None, var=None):
    """extract_structure_observables_from_out (local_structure.py:55-103) for every replica of a batch.
    Returns a dict of device tensors with a leading replica axis."""
    if amp is None or var is None:
        amp, _, var = fft_amplitudes(rb)
    M = rb.M
    s = int(start_fraction * M)
    if k_max is not None:
        amp = amp[:, :, :k_max]
    fft_mean = amp[:, s:].mean(dim=1)
    fft_std = amp[:, s:].std(dim=1, unbiased=True)
    k_cut = min(25, fft_mean.shape[1])
    out = dict(var_mean=var[:, s:].mean(dim=1), var_std=var[:, s:].std(dim=1, unbiased=True), fft_mean=fft_mean,
               fft_std=fft_std, dominant_k=fft_mean[:, 1:].argmax(dim=1) + 1, low_k_power=fft_mean[:, 1:k_cut].sum(dim=1),
               lowk_variance=(amp[:, s:, 1:k_cut] ** 2).sum(dim=2).mean(dim=1))
    if rb.obs_m_local is not None:
        ml = rb.obs_m_local[:, s:].reshape(rb.R, -1)
        out["m_local_var"] = ml.var(dim=1, unbiased=False)
    return out


def time_to_pattern(amp, times_obs, threshold=0.05, k=1):
    """First observation time at which mode k exceeds `threshold`, NaN if it never does (:195-202).  amp: [R][M][k]."""
    t = torch.as_tensor(times_obs, dtype=torch.float64, device=amp.device)
    hit = amp[:, :, k] > threshold
    first = torch.where(hit.any(dim=1), hit.to(torch.int8).argmax(dim=1), torch.full((amp.shape[0],), -1, device=amp.device))
    out = torch.full((amp.shape[0],), float("nan"), dtype=torch.float64, device=amp.device)
    ok = first >= 0
    out[ok] = t[first[ok]]
    return out


def ensemble_time_to_pattern(ttp):
    """Mean and standard error over the runs that formed a pattern (:203-209; np.std, i.e. ddof = 0)."""
    v = ttp[~torch.isnan(ttp)]
    if v.numel() == 0:
        return float("nan"), float("nan")
    return float(v.mean()), float(v.std(unbiased=False) / np.sqrt(v.numel()))


def temporal_autocorrelation(total, lag=1):
    """mean_t mean_x total[t] * total[t + lag] (:223-231).  total: [R][M][L] -> [R]."""
    M = total.shape[1]
    if M - lag <= 0:
        return torch.full((total.shape[0],), float("nan"), dtype=torch.float64, device=total.device)
    return (total[:, : M - lag] * total[:, lag:]).mean(dim=2).mean(dim=1)


def lowk_variance_time(amp, k_cut=25):
    """sum_{k=1..k_cut} |rho_hat_k|^2 per observation row (:232-234).  amp: [R][M][k] -> [R][M]."""
    return (amp[:, :, 1:k_cut + 1] ** 2).sum(dim=2)


def _np_or_torch(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))


def spectral_entropy(fft_mean, k_max=None):
    """-sum p log(p + 1e-12), p = normalised power of the modes k >= 1 (:235-241).  fft_mean: [..., k]."""
    f = _np_or_torch(fft_mean)
    if k_max is not None:
        f = f[..., :k_max]
    power = f[..., 1:] ** 2
    p = power / power.sum(dim=-1, keepdim=True)
    return -(p * torch.log(p + 1e-12)).sum(dim=-1)


def mode_competition_ratio(fft_mean):
    """Largest mode amplitude over the sum of the others (:242-245).  fft_mean: [..., k]."""
    a = _np_or_torch(fft_mean)[..., 1:]
    top = a.max(dim=-1).values
    return top / (a.sum(dim=-1) - top + 1e-12)


def extract_growth_rate(amp, times_obs, k=1, t_min=0.0, t_max=None, amp_min=1e-4):
    """Slope of the least-squares line through (t, log amp_k) over the rows with t_min <= t (<= t_max) and
    amp_k > amp_min; NaN when fewer than 3 rows qualify (:246-265, np.polyfit degree 1).  amp: [R][M][k] -> [R]."""
    t = torch.as_tensor(times_obs, dtype=torch.float64, device=amp.device)[None, :]
    a = amp[:, :, k]
    mask = (t >= t_min) & (a > amp_min)
    if t_max is not None:
        mask = mask & (t <= t_max)
    w = mask.to(torch.float64)
    cnt = w.sum(dim=1)
    y = torch.log(torch.where(mask, a, torch.ones_like(a)))
    safe = cnt.clamp(min=1.0)
    tb = (w * t).sum(dim=1) / safe
    yb = (w * y).sum(dim=1) / safe
    dt = (t - tb[:, None]) * w
    den = (dt * dt).sum(dim=1)
    slope = (dt * (y - yb[:, None])).sum(dim=1) / torch.where(den > 0, den, torch.ones_like(den))
    return torch.where((cnt >= 3) & (den > 0), slope, torch.full_like(slope, float("nan")))


def cluster_size_distribution(rho, threshold):
    """Lengths of the maximal runs of sites with rho > threshold, in lattice order (:210-222).  One profile, host."""
    occ = np.asarray(rho) > threshold
    edge = np.diff(np.concatenate([[0], occ.astype(np.int8), [0]]))
    return (np.flatnonzero(edge == -1) - np.flatnonzero(edge == 1)).astype(int)
