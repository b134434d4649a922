"""CPU restatement (numpy) of the deterministic part of the reference's IMEX PDE solver — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(aps_b200.imex_pde -> libaps_b200.so) never does.

Follows IMEX_PDE_solver_class.py: magnetization() :157-169, advective_derivative() :171-185, step() :187-233 and the
per-step bookkeeping of solve() :241-254, written independently of the reference's formulation so that it is a
check and not a copy: the implicit diffusion is solved spectrally (periodic: the circulant eigenvalues
1 + 2a(1 - cos(2 pi k / L))) or with a banded solve (Neumann), the kernel convolution is a direct ring sum.
Pinned against short runs of the unmodified reference (tests/golden/pde_*.npz, tools/gen_golden.py): agreement to
rounding (tests state 1e-10), not bitwise.  Tracers are not restated (they consume numpy's global random stream).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_banded


def ring_kernel(L, dx, kernel_sigma):
    """_build_kernel (:87-96): Gaussian of the ring distance, normalised."""
    j = np.arange(L)
    k = np.exp(-0.5 * (np.minimum(j, L - j) * dx / kernel_sigma) ** 2)
    return k / k.sum()


def ring_convolve(x, kernel):
    L = x.size
    out = kernel[0] * x
    for d in range(1, L):
        out = out + kernel[d] * np.roll(x, d)          # out[i] += kernel[d] * x[i - d]; kernel is symmetric on the ring
    return out


def local_m(rho_p, rho_m, kernel):
    if kernel is None:
        return (rho_p - rho_m) / (rho_p + rho_m + 1e-12)
    return ring_convolve(rho_p - rho_m, kernel) / (ring_convolve(rho_p + rho_m, kernel) + 1e-12)


def diffuse(x, a, bc):
    if a == 0.0:
        return x.copy()
    L = x.size
    if bc == "periodic":
        lam = 1.0 + 2.0 * a * (1.0 - np.cos(2.0 * np.pi * np.arange(L // 2 + 1) / L))
        return np.fft.irfft(np.fft.rfft(x) / lam, n=L)
    ab = np.zeros((3, L))
    ab[0, 1:] = -a; ab[1, :] = 1.0 + 2.0 * a; ab[2, :-1] = -a
    ab[0, 1] = -2.0 * a; ab[2, L - 2] = -2.0 * a          # rows 0 and L-1 of the Neumann matrix (:79-81)
    return solve_banded((1, 1), ab, x)


def upwind(rho, direction, dx, bc):
    d = np.zeros_like(rho)
    if direction > 0:
        d[1:] = np.diff(rho) / dx
        d[0] = 0.0 if bc == "neumann" else (rho[0] - rho[-1]) / dx
    else:
        d[:-1] = np.diff(rho) / dx
        d[-1] = 0.0 if bc == "neumann" else (rho[0] - rho[-1]) / dx
    return d


def cw(beta, sigma, m):
    return np.clip(np.exp(-beta * sigma * m), 1e-8, 1e8)


def run(L, xlim, T, dt, gamma, lam, beta, bc, active_model, gaussian_kernel, kernel_sigma, snapshot_interval,
        rho_p0, rho_m0, **_ignored):
    dx = xlim / L
    nsteps = int(T / dt)
    a = gamma * dt / dx ** 2
    kernel = ring_kernel(L, dx, kernel_sigma) if gaussian_kernel else None
    p, m = np.array(rho_p0, dtype=float), np.array(rho_m0, dtype=float)
    m_series, var_series = np.zeros(nsteps + 1), np.zeros(nsteps + 1)
    snaps, msnaps, times = [], [], []
    for n in range(nsteps + 1):
        mf = local_m(p, m, kernel)
        m_series[n] = mf.mean()
        var_series[n] = np.var(p + m)
        if n % snapshot_interval == 0:
            snaps.append(p + m); msnaps.append(p - m); times.append(n * dt)
        if n == nsteps:
            break
        dp, dm = diffuse(p, a, bc), diffuse(m, a, bc)
        react = cw(beta, -1, mf) * dm - cw(beta, +1, mf) * dp
        if active_model == "bidirectional":
            new_p = np.clip(dp + dt * (-lam * upwind(dp, +1, dx, bc) + react), 0, None)
            new_m = np.clip(dm + dt * (+lam * upwind(dm, -1, dx, bc) - react), 0, None)
        else:
            p_star = np.clip(dp + dt * react, 0, None)
            new_m = np.clip(dm - dt * react, 0, None)
            new_p = np.clip(p_star + dt * (-lam * upwind(p_star, +1, dx, bc)), 0, None)
        scale = (dp + dm).sum() / (new_p + new_m).sum()
        p, m = new_p * scale, new_m * scale
    return dict(rho_p=p, rho_m=m, m_series=m_series, var_series=var_series, snapshots=np.array(snaps),
                m_snapshots=np.array(msnaps), times=np.array(times))
