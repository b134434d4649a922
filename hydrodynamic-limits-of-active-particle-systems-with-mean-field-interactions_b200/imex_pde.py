"""`IMEXPDE` — host-side mirror of the reference's hydrodynamic PDE solver (IMEX_PDE_solver_class.py:11-307) on top of
the batched CUDA stepper (include/aps_pde.h, csrc/aps_pde.cu).  Same constructor keywords, `initialize()`, `solve()`,
`get_output()` keys, so the unchanged run / sweep scripts (`IMEX_PDE_solver_run*.py`) can import it instead.

What is identical to the reference for a given `seed`: the initial fields and tracers (`initialize()` consumes a
`numpy.random.RandomState(seed)` in the reference's call order, which is the stream `np.random.seed(seed)` gives it),
and — to rounding, see tests — every deterministic output of `solve()`: final `rho_p`/`rho_m`, `m_series`,
`var_series`, `snapshots`, `m_snapshots`, `times`.  What is equal in distribution only: `v_eff_series` / `D_eff_series`
(the reference draws the tracer noise from numpy's global stream inside the time loop; the kernel uses Philox).
`fft_amp` / `fft_phase`: `solve()` returns them for EVERY step, shape (nsteps+1, L/2+1) like the reference (:246-249;
the kernel streams `rho_p + rho_m` of every step to HBM and one batched cuFFT transforms the rows — 320 MB per run at
the reference's default sizes, as in the reference).  `solve_many(..., spectra="snapshots")`, the default for batches,
returns them for the snapshot rows only (`fft_times`): a sweep of 33 runs x 80 000 steps would need 42 GB otherwise.
`plot_all()` / `plot_individual()` are the reference's own plotting code, borrowed from its file (see `reference_class`).

`solve_many(solvers)` runs any number of compatible instances (same L, dt, T, bc, model, kernel mode, tracer count)
in ONE launch, one CTA per instance — this is how the sweep scripts' `for beta: for run:` loops map onto the GPU.
There is no CPU fallback: without the CUDA library / an sm_100 device `solve()` raises.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import capi


class IMEXPDE:
    def __init__(self, L=1000, xlim=1.0, T=10.0, dt=5e-4, gamma=2.33e-4, lam=0.6, beta=2.0, bc="periodic",
                 active_model="bidirectional", gaussian_kernel=False, kernel_sigma=0.02, snapshot_interval=50,
                 outdir="IMEX_output", seed=None):
        self.L = L
        self.xlim = xlim
        self.dx = xlim / L
        self.x = np.linspace(0, xlim, L, endpoint=False)
        self.T = T
        self.dt = dt
        self.nsteps = int(T / dt)
        self.gamma = gamma
        self.lam = lam
        self.beta = beta
        if bc not in capi.APS_PDE_BC:
            raise ValueError(f"unknown bc {bc!r}")
        if active_model not in capi.APS_PDE_MODEL:
            raise ValueError(f"unknown active_model {active_model!r}")
        self.bc = bc
        self.active_model = active_model
        self.gaussian_kernel = gaussian_kernel
        self.kernel_sigma = kernel_sigma
        if gaussian_kernel and kernel_sigma > 100000:
            # the reference's scalar-magnetisation branch (:162-165) cannot be indexed by its own tracer code (:257)
            raise NotImplementedError("kernel_sigma > 1e5 crashes the reference's solve(); use a value below 1e5")
        self.snapshot_interval = snapshot_interval
        self.seed = seed
        self.outdir = Path(outdir)
        self.outdir.mkdir(exist_ok=True)                        # :53
        self._rs = np.random.RandomState(seed) if seed is not None else np.random.RandomState()
        self.rho_mean = 1.0 / self.xlim
        self.kernel = None
        self.kernel_radius = 0
        if gaussian_kernel:                                    # _build_kernel (:87-96)
            i = np.arange(L)
            k = np.exp(-0.5 * (np.minimum(i, L - i) * self.dx / kernel_sigma) ** 2)
            self.kernel = k / k.sum()
            half = (L - 1) // 2
            self.kernel_radius = next((r for r in range(half + 1) if self.kernel[r + 1:L - r].sum() <= 1e-22), L)
        self._out = None

    # ---- initial condition: same numpy call sequence as initialize() (:99-137) ----
    def initialize(self, mode="poisson", rho0=1.0, noise=0.2, n_tracers=1000):
        rs = self._rs
        if mode == "homogeneous":
            rho_p = rho0 + noise * rs.randn(self.L)
            rho_m = rho0 + noise * rs.randn(self.L)
        elif mode == "poisson":
            rho_p = np.exp(-np.abs(self.x - 0.5) / 0.05)
            rho_m = np.exp(-np.abs(self.x - 0.5) / 0.05)
            rho_p += noise * rs.randn(self.L)
            rho_m += noise * rs.randn(self.L)
        else:
            raise ValueError("Unknown init mode.")
        rho_p = np.clip(rho_p, 0, None)
        rho_m = np.clip(rho_m, 0, None)
        tot = (rho_p + rho_m).sum()
        self.rho_p = rho_p / tot
        self.rho_m = rho_m / tot
        self.n_tracers = n_tracers
        self.tracers = rs.choice(self.L, size=n_tracers) * self.dx
        self.tracers_unwrapped = self.tracers.copy()
        self.tracer_state = rs.choice([-1, 1], size=n_tracers)
        self._out = None

    def solve(self):
        solve_many([self], spectra="all")

    # ---- plotting: the reference's own code (IMEX_PDE_solver_class.py:309-461), applied to this instance ----
    def plot_all(self, *a, **k):
        return reference_class().plot_all(self, *a, **k)

    def plot_individual(self, *a, **k):
        return reference_class().plot_individual(self, *a, **k)

    def get_output(self):
        if self._out is None:
            raise RuntimeError("call solve() first")
        return dict(self._out)


_REF_CLASS = None


def reference_class():
    """The reference's `IMEXPDE` loaded under a private module name from $APS_REFERENCE_PATH, <repo>/baseline/_ref or
    /root/reference; only its plot methods are used (they read the attributes solve() fills in)."""
    global _REF_CLASS
    if _REF_CLASS is None:
        import importlib.util
        import os
        for d in (os.environ.get("APS_REFERENCE_PATH"), os.path.join(capi.REPO_ROOT, "baseline", "_ref"), "/root/reference"):
            f = os.path.join(d, "IMEX_PDE_solver_class.py") if d else None
            if f and os.path.exists(f):
                spec = importlib.util.spec_from_file_location("_aps_reference_IMEX_PDE_solver_class", f)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)                    # needs matplotlib, like the reference
                _REF_CLASS = mod.IMEXPDE
                break
        else:
            raise RuntimeError("plot_all / plot_individual are delegated to the reference's code: set APS_REFERENCE_PATH "
                               "to a directory holding IMEX_PDE_solver_class.py (or run tools/install_reference.py)")
    return _REF_CLASS


def _same(solvers, attr):
    vals = {getattr(s, attr) for s in solvers}
    if len(vals) != 1:
        raise ValueError(f"solve_many: all instances must share `{attr}` (got {sorted(map(str, vals))})")
    return vals.pop()


def solve_many(solvers, device=None, spectra="snapshots"):
    """Run every instance of `solvers` through its nsteps steps in one kernel launch (one CTA per instance) and
    attach the reference's output dict to each (`get_output()`).  spectra="all": fft_amp / fft_phase for every step
    (the reference's shapes); "snapshots": for the snapshot rows only."""
    import torch

    if not torch.cuda.is_available():
        raise capi.ApsError("no CUDA device: the IMEX PDE stepper has no CPU path")
    lib = capi.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    L, nsteps, dt, xlim = (_same(solvers, a) for a in ("L", "nsteps", "dt", "xlim"))
    bc, model, gk, ntr, interval = (_same(solvers, a) for a in ("bc", "active_model", "gaussian_kernel", "n_tracers",
                                                                 "snapshot_interval"))
    R = len(solvers)
    f64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    beta, lam, gamma = (f64([getattr(s, a) for s in solvers]) for a in ("beta", "lam", "gamma"))
    rho_p, rho_m = f64(np.stack([s.rho_p for s in solvers])), f64(np.stack([s.rho_m for s in solvers]))
    kernel = f64(np.stack([s.kernel for s in solvers])) if gk else None
    radius = torch.as_tensor(np.array([s.kernel_radius for s in solvers], np.int32)).to(dev) if gk else None
    n_snap = nsteps // interval + 1
    z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
    m_series, var_series = z(R, nsteps + 1), z(R, nsteps + 1)
    snaps, msnaps = z(R, n_snap, L), z(R, n_snap, L)
    window = max(1, int(0.05 / dt))                            # :237-238
    seeds = torch.as_tensor(np.array([(s.seed if s.seed is not None else 0) & (2 ** 63 - 1) for s in solvers], np.int64)).to(dev)
    if ntr > 0:
        tpos = f64(np.stack([s.tracers_unwrapped for s in solvers]))
        tstate = torch.as_tensor(np.stack([np.asarray(s.tracer_state, np.int8) for s in solvers])).to(dev)
        hist = z(R, window, ntr)
        v_eff, d_eff = z(R, nsteps + 1), z(R, nsteps + 1)
    ptr = lambda t: None if t is None else t.data_ptr()
    args = capi.ApsPdeArgs(L=L, n_runs=R, bc=capi.APS_PDE_BC[bc], model=capi.APS_PDE_MODEL[model], field=1 if gk else 0,
                           snapshot_interval=interval, n_tracers=ntr, window=window, nsteps=nsteps, dt=dt, dx=xlim / L,
                           xlim=xlim, beta=ptr(beta), lam=ptr(lam), gamma=ptr(gamma), kernel=ptr(kernel), radius=ptr(radius),
                           seeds=ptr(seeds), rho_p=ptr(rho_p), rho_m=ptr(rho_m), m_series=ptr(m_series),
                           var_series=ptr(var_series), snapshots=ptr(snaps), m_snapshots=ptr(msnaps),
                           tracer_pos=ptr(tpos) if ntr else None, tracer_state=ptr(tstate) if ntr else None,
                           tracer_hist=ptr(hist) if ntr else None, v_eff_series=ptr(v_eff) if ntr else None,
                           D_eff_series=ptr(d_eff) if ntr else None)
    tot_series = z(R, nsteps + 1, L) if spectra == "all" else None
    args.tot_series = ptr(tot_series)
    capi.check(lib.aps_pde_solve_device(args, torch.cuda.current_stream().cuda_stream), "aps_pde_solve_device")
    if spectra == "all":                                       # :247-249, every step
        fft = torch.cat([torch.fft.rfft(tot_series[:, i:i + 8192], dim=-1) / L for i in range(0, nsteps + 1, 8192)], dim=1)
        del tot_series
    else:
        fft = torch.fft.rfft(snaps, dim=-1) / L                # spectra of the snapshot rows only
    h = lambda t: t.cpu().numpy()
    rho_p_h, rho_m_h, m_h, var_h, snaps_h, msnaps_h, fft_h = map(h, (rho_p, rho_m, m_series, var_series, snaps, msnaps, fft))
    times = np.arange(n_snap) * interval * dt
    nan_series = np.full(nsteps + 1, np.nan)
    for r, s in enumerate(solvers):
        s.rho_p, s.rho_m = rho_p_h[r], rho_m_h[r]
        s.m_series, s.var_series = m_h[r], var_h[r]
        if ntr > 0:
            s.tracers_unwrapped = tpos[r].cpu().numpy()
            s.tracers = s.tracers_unwrapped % s.xlim
            s.tracer_state = tstate[r].cpu().numpy().astype(int)
        s.fft_amp, s.fft_phase = np.abs(fft_h[r]), fft_h[r]    # attributes the reference's plot methods read
        s.snapshots, s.m_snapshots, s.times = list(snaps_h[r]), list(msnaps_h[r]), list(times)
        s.v_eff_series = v_eff[r].cpu().numpy() if ntr > 0 else nan_series.copy()
        s.D_eff_series = d_eff[r].cpu().numpy() if ntr > 0 else nan_series.copy()
        s._out = dict(rho_p=s.rho_p, rho_m=s.rho_m, m_series=s.m_series, var_series=s.var_series,
                      fft_amp=np.abs(fft_h[r]), fft_phase=fft_h[r], fft_times=times if spectra != "all" else np.arange(nsteps + 1) * dt, snapshots=snaps_h[r],
                      m_snapshots=msnaps_h[r], times=times, v_eff_series=s.v_eff_series, D_eff_series=s.D_eff_series)
    return solvers
