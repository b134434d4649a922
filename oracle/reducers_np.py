"""numpy restatement of the sweep drivers' per-run reducers (TEST INFRASTRUCTURE ONLY).

Follows PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py:
  v_eff_and_window       :123-162   (incl. the `~safe[start_idx:]` index-array quirk, :141-154)
  rho_eff                :165-194
  blocking_probability   :197-229
  mean_magnetisation     :316-319
  D_eff_active           :500-525
Pinned against tests/golden/reducers.json, which holds the outputs of the reference's own
functions (extracted by AST in tools/gen_golden.py) on two recorded runs.
"""
import numpy as np


def v_eff_and_window(times, total, L, boundary_xmin=0.99, max_frac=0.06, min_window_fraction=0.10):
    M = total.shape[0]
    x_grid = np.linspace(0, 1.0, L)
    dx = x_grid[1] - x_grid[0]
    boundary = total[:, x_grid >= boundary_xmin].sum(axis=1) * dx
    N_t = total.sum(axis=1) * dx
    frac = boundary / (N_t + 1e-12)
    unsafe_idx = np.where(frac >= max_frac)[0]
    start = int(0.65 * M)
    end = M
    if unsafe_idx.size and len(unsafe_idx[start:]) > 0:      # `np.where(~idx)` is all-true for index arrays
        end = start
        min_len = max(3, int(min_window_fraction * M))
        if end - start < min_len:
            end = min(M, start + min_len)
    mean_x = (total * x_grid).sum(axis=1) / (total.sum(axis=1) + 1e-12)
    v_eff = np.gradient(mean_x, times)
    return float(np.mean(v_eff[start:end])), v_eff, start, end, frac


def rho_eff(total, start, end, window_fraction=0.05):
    M, L = total.shape
    x_grid = np.linspace(0, 1.0, L)
    dx = x_grid[1] - x_grid[0]
    vals = []
    for t in range(start, end):
        occ = np.where(total[t] > 0)[0]
        if len(occ) == 0:
            continue
        x_max = x_grid[occ[-1]]
        mask = (x_grid >= x_max - window_fraction) & (x_grid <= x_max)
        if mask.sum() == 0:
            continue
        vals.append(total[t][mask].sum() * dx / window_fraction)
    return float(np.mean(vals))


def blocking_probability(total, rho_p, start, end):
    M, L = total.shape
    blocked = attempts = 0.0
    for t in range(start, end):
        for i in np.where(rho_p[t] > 0)[0]:
            if i + 1 >= L:
                continue
            attempts += rho_p[t][i]
            if total[t][i + 1] >= 1.0:
                blocked += rho_p[t][i]
    return 0.0 if attempts == 0 else blocked / attempts


def mean_magnetisation(m_global, start, end):
    return float(np.mean(np.asarray(m_global, dtype=float)[start:end]))


def d_eff_active(times, pos_list, dx, start, end):
    S, tv = [], []
    pos0 = pos_list[start] * dx
    for k in range(start + 1, end):
        pos_t = pos_list[k] * dx
        n = min(len(pos0), len(pos_t))
        if n < 2:
            continue
        ri = pos_t[:n] - pos0[:n]
        S.append(np.sum((ri - np.mean(ri)) ** 2) / (n - 1))
        tv.append(times[k] - times[start])
    return float(np.polyfit(tv, S, 1)[0])


def all_reducers(times, rho_p, rho_m, total, m_global, pos_list, L, dx):
    mean_v, v_eff, si, ei, frac = v_eff_and_window(times, total, L)
    return dict(mean_v=mean_v, si=si, ei=ei, D_eff=d_eff_active(times, pos_list, dx, si, ei),
                m_mean=mean_magnetisation(m_global, si, ei), rho_eff=rho_eff(total, si, ei),
                block=blocking_probability(total, rho_p, si, ei), v_eff=v_eff, frac_boundary=frac)
