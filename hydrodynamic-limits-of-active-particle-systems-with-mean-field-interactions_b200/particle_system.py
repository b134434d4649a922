"""Host-side mirror of the reference's `ParticleSystem` (PARTICLE_solver_CLASS.py:13-558).

Same constructor keywords, same entry points (`run`, `step_gillespie`, `init_particles`,
`compute_local_m_field`, `empirical_densities_from_particles`, `_build_occupancy`), same
returned dict — but the time-stepping loop runs in the K1 CUDA kernel through the C ABI.

Random numbers, two modes selected by the `rng=` argument (the reference's injection seam, :26,75-78):
  * a numpy `Generator` (or any object with exponential/random/choice/poisson): REPLAY mode.
    The variates the reference would have drawn are drawn from the same object in the same
    order and injected into the kernel, so `ParticleSystem(rng=np.random.default_rng(s)).run()`
    reproduces the reference's trajectory for seed `s` bit for bit.  The conditional 4th draw of a
    diffusive hop (:378) is handled by speculative chunks with rewind of the bit generator.
  * `PhiloxRNG(seed)` or `rng=None`: NATIVE mode, in-kernel counter-based Philox4x32-10.

Anchors / binding / unbinding / exit (`anchor_positions`, `k_on`, `k_off`, `k_exit`, CLASS.py:307-348,418-436) are supported
(generic K1 kernel).  `periodic=True` (ring: hops wrap, :278-288) is supported too; its local field is the reference's ring
kernel (:111-121) applied as a truncated direct sum instead of the FFT of :224-227, so trajectories match the reference for
a given seed and `m_local_list` agrees to 1e-13 rather than bit for bit.
A custom `flip_rate_fn(sigma, m)` (CLASS.py:59-62) is evaluated on the host on a grid of 8193 magnetisation values per
orientation and interpolated linearly in the kernel (`engine.tabulate_flip_rate`, `aps_flip_interp`): rates agree with the
callable to (2/8192)^2 max|f''|/8, so a trajectory equals the reference's for a given seed unless a uniform variate falls
within that relative distance of a decision threshold (the fixture `custom_flip_*` matches event for event).
"""
from __future__ import annotations

import os

import numpy as np

from . import capi
from .capi import APS_REC_COUNTS, APS_REC_MLOCAL, APS_REC_POS


class PhiloxRNG:
    """Selects the native in-kernel Philox stream.  Initial conditions (host side, as in the
    reference) are drawn from a numpy Generator seeded with the same seed."""

    def __init__(self, seed=None):
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._gen = np.random.default_rng(self.seed)

    def choice(self, *a, **k):
        return self._gen.choice(*a, **k)

    def poisson(self, *a, **k):
        return self._gen.poisson(*a, **k)

    # single-step entry (`step_gillespie`) draws its variates on the host like the reference does (CLASS.py:358-362,378)
    def exponential(self, *a, **k):
        return self._gen.exponential(*a, **k)

    def random(self, *a, **k):
        return self._gen.random(*a, **k)


_REF_CLASS = None


def reference_class():
    """The reference's own `ParticleSystem` (PARTICLE_solver_CLASS.py), loaded under a private module name from
    $APS_REFERENCE_PATH, <repo>/baseline/_ref (tools/install_reference.py) or /root/reference.  Only the plotting /
    animation methods (CLASS.py:561-1093, out of scope for the CUDA path: they post-process the `out` dict) are
    borrowed from it; it needs matplotlib and vispy like the reference itself."""
    global _REF_CLASS
    if _REF_CLASS is not None:
        return _REF_CLASS
    import importlib.util
    cands = [os.environ.get("APS_REFERENCE_PATH"), os.path.join(capi.REPO_ROOT, "baseline", "_ref"), "/root/reference"]
    for d in cands:
        f = os.path.join(d, "PARTICLE_solver_CLASS.py") if d else None
        if f and os.path.exists(f):
            spec = importlib.util.spec_from_file_location("_aps_reference_PARTICLE_solver_CLASS", f)
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except ImportError as e:
                raise ImportError(f"the reference's plotting code at {f} could not be imported ({e}); it needs "
                                  "matplotlib and vispy exactly like the reference does") from e
            _REF_CLASS = mod.ParticleSystem
            return _REF_CLASS
    raise RuntimeError("visualize_all / plot_individuals / animate_profiles are the reference's own post-processing of "
                       "the `out` dict and are delegated to its code: set APS_REFERENCE_PATH to a directory holding the "
                       "reference's PARTICLE_solver_CLASS.py (or run tools/install_reference.py)")


class ParticleSystem:
    def __init__(self, L, xlim, rate_diffusion, rate_active, beta, flip_rate_fn=None, init="fixed", N=1000,
                 rho0_plus=None, rho0_minus=None, rng=None, scale_rates=True, local_kernel_sigma=0.005,
                 periodic=False, minus_anchor=True, immobilize_when_anchored=True, anchor_positions=None,
                 anchor_radius=0.005, site_capacity=1, crowding_suppresses_rates=False, k_on=0.1, k_off=0.01,
                 suppress_flip_when_bound=True, k_exit=0):
        self.L = L
        self.xlim = xlim
        self.K = site_capacity
        self.dx = self.xlim / self.L
        if scale_rates:                                   # CLASS.py:45-50
            self.rate_diffusion = rate_diffusion / (self.dx ** 2)
            self.rate_active = rate_active / (self.dx)
        else:
            self.rate_diffusion = float(rate_diffusion)
            self.rate_active = float(rate_active)
        self.beta = beta
        self.k_on, self.k_off, self.k_exit = k_on, k_off, k_exit
        self.suppress_flip_when_bound = suppress_flip_when_bound
        self.crowding_suppresses_rates = crowding_suppresses_rates
        if flip_rate_fn is None:                          # CLASS.py:59-62
            self.flip_rate_fn = lambda sigma, m: np.exp(-self.beta * sigma * m)
            self._flip_tab = None                         # evaluated in the kernel (exp with single-rounding operations)
        else:
            from .engine import tabulate_flip_rate
            self.flip_rate_fn = flip_rate_fn
            self._flip_tab = tabulate_flip_rate(flip_rate_fn)     # host-evaluated table, interpolated on the device
        assert init in ("fixed", "poisson")
        self.init_mode = init
        if self.init_mode == "fixed":
            self.N_fixed = N
        else:
            self.rho0_plus = np.array([rho0_plus(i / self.L) for i in range(self.L)], dtype=float)
            self.rho0_minus = np.array([rho0_minus(i / self.L) for i in range(self.L)], dtype=float)
        self.rng = PhiloxRNG() if rng is None else rng
        self.local_kernel_sigma = local_kernel_sigma
        self.periodic = periodic
        self.immobilize_when_anchored = immobilize_when_anchored
        self.minus_anchor = minus_anchor
        self._sigma_grid = self.local_kernel_sigma / self.dx
        self.anchor_radius = anchor_radius
        # anchor sites: positions -> lattice indices within anchor_radius (CLASS.py:88-104)
        self.is_anchor_site = np.zeros(self.L, dtype=bool)
        self.anchor_idx_array = np.array([], dtype=int)
        if anchor_positions is None:
            self.anchor_positions = None
            self.anchor_idxs = np.array([], dtype=int)
        else:
            self.anchor_positions = anchor_positions
            ap = np.asarray(anchor_positions, dtype=float)
            self.anchor_idxs = np.unique(np.round((ap / self.xlim) * (self.L - 1)).astype(int))
            r_idx = int(np.ceil(anchor_radius / self.dx))
            sites = set()
            for a in self.anchor_idxs:
                sites.update(range(max(0, a - r_idx), min(self.L - 1, a + r_idx) + 1))
            self.anchor_idx_array = np.array(sorted(sites), dtype=int)
            self.is_anchor_site[self.anchor_idx_array] = True
        self._has_anchors = anchor_positions is not None
        self._kernel = None
        self._fft_kernel = None
        from .engine import gaussian_weights, periodic_weights
        if self.local_kernel_sigma > 0 and self.periodic:
            # ring kernel of CLASS.py:111-121, applied as a truncated direct sum instead of the FFT of :224-227
            self._radius, self._weights = periodic_weights(self.L, self.dx, self.local_kernel_sigma)
        elif self.local_kernel_sigma > 0:
            self._radius, self._weights = gaussian_weights(self._sigma_grid)
        else:
            self._radius, self._weights = -1, np.zeros(1)
        self.last_run_info = None

    # ---- initial conditions: host side, same rng call sequence as CLASS.py:141-195 ----------
    def _init_fixed(self):
        N = self.N_fixed
        if self.K == 1:
            pos = self.rng.choice(self.L, size=N, replace=False)
        else:
            pos = np.empty(N, dtype=np.int64)
            fill = np.zeros(self.L, dtype=int)
            for i in range(N):
                site = self.rng.choice(np.where(fill < self.K)[0])
                pos[i] = site
                fill[site] += 1
        sigma = self.rng.choice([1, -1], size=N)
        return pos.astype(np.int64), sigma.astype(np.int8)

    def _init_poisson(self):
        cp = self.rng.poisson(self.rho0_plus)
        cm = self.rng.poisson(self.rho0_minus)
        pos, sig = [], []
        for x in np.nonzero(cp + cm)[0]:
            labels = np.array([1] * int(cp[x]) + [-1] * int(cm[x]), dtype=int)
            if labels.size > self.K:
                labels = labels[self.rng.choice(labels.size, size=self.K, replace=False)]
            pos.extend([x] * labels.size)
            sig.extend(labels.tolist())
        return np.asarray(pos, dtype=np.int64), np.asarray(sig, dtype=np.int8)

    def init_particles(self):
        return self._init_fixed() if self.init_mode == "fixed" else self._init_poisson()

    @staticmethod
    def empirical_densities_from_particles(pos, sigma, L, dx, total_norm=None):
        """API-parity helper (CLASS.py:198-214); run() itself uses the device expansion kernel."""
        counts_p = np.bincount(pos[sigma == 1], minlength=L)
        counts_m = np.bincount(pos[sigma == -1], minlength=L)
        denom = float(max(1, pos.size)) * dx if total_norm is None else float(total_norm) * dx
        return (counts_p / denom).astype(float), (counts_m / denom).astype(float)

    def _build_occupancy(self, pos, sigma):
        counts_p = np.bincount(pos[sigma == 1], minlength=self.L).astype(int)
        counts_m = np.bincount(pos[sigma == -1], minlength=self.L).astype(int)
        return counts_p + counts_m, counts_p, counts_m

    def compute_local_m_field(self, counts_p, counts_m):
        """CLASS.py:216-246 evaluated by the device field kernel through the C ABI."""
        lib = capi.load()
        cp = np.ascontiguousarray(counts_p, dtype=np.int32)
        cm = np.ascontiguousarray(counts_m, dtype=np.int32)
        out = np.zeros(self.L, dtype=np.float64)
        from .batch import make_params
        p = make_params(self.L, self.K, self._radius, self.rate_diffusion, self.rate_active, 0.0,
                        capi.APS_FLAG_PERIODIC if self.periodic else 0)
        w = np.ascontiguousarray(self._weights, dtype=np.float64)
        capi.check(lib.aps_m_field_host(p, w.ctypes.data, cp.ctypes.data, cm.ctypes.data, out.ctypes.data),
                   "aps_m_field_host")
        return out

    # ---- the hot path ----------------------------------------------------------------------
    def _make_batch(self, pos, sigma, times_obs, T, record, seeds=None):
        from .engine import ReplicaBatch
        n = int(pos.size)
        return ReplicaBatch(L=self.L, K=self.K, radius=self._radius, weights=self._weights, D=self.rate_diffusion,
                            lam=self.rate_active, T=T, times_obs=times_obs, betas=[float(self.beta)], n=[n],
                            pos0=np.asarray(pos, dtype=np.int32).reshape(1, -1) if n else np.zeros((1, 1), np.int32),
                            sigma0=np.asarray(sigma, dtype=np.int8).reshape(1, -1) if n else np.ones((1, 1), np.int8),
                            seeds=seeds, record=record, crowding=self.crowding_suppresses_rates, dx=self.dx,
                            anchor_mask=self.is_anchor_site if self._has_anchors else None, k_on=self.k_on, k_off=self.k_off,
                            k_exit=self.k_exit, suppress_flip_when_bound=self.suppress_flip_when_bound,
                            immobilize_when_anchored=self.immobilize_when_anchored, periodic=self.periodic,
                            flip_tab=self._flip_tab)

    def run(self, T=10.0, obs_dt=0.01, record_fft=False, record_var=False):
        import torch

        pos, sigma = self.init_particles()
        n = int(pos.size)
        times_obs = np.arange(0.0, T, obs_dt)
        M = len(times_obs)
        if n == 0:
            # the reference fails here too (ZeroDivisionError / tuple-unpack ValueError, CLASS.py:220,257,513)
            raise ValueError("ParticleSystem.run: no particles (the reference raises at this point)")
        native = isinstance(self.rng, PhiloxRNG)
        rb = self._make_batch(pos, sigma, times_obs, T, APS_REC_COUNTS | APS_REC_POS | APS_REC_MLOCAL,
                              seeds=[self.rng.seed] if native else None)
        if native:
            rb.run_philox()
            launches = 1
        else:
            launches = self._run_replay(rb)
        status = int(rb.status.item())
        if status == capi.APS_RUN_EMPTY:
            raise ValueError("total rate R <= 0 (the reference raises at this point, CLASS.py:355,513)")
        n_obs = int(rb.n_obs.item())
        rho_p, rho_m, total, var = rb.expand(want_var=record_fft and record_var)
        out_pos = rb.obs_pos[0].cpu().numpy()
        sig_sum = rb.obs_sigma_sum[0].cpu().numpy()
        counts = rb.obs_n[0].cpu().numpy()                        # particles per row (exits shrink the system)
        bounds = rb.obs_bound[0].cpu().numpy().astype(bool) if rb.obs_bound is not None else None
        m_global = np.zeros(M, dtype=float)
        m_global[:n_obs] = sig_sum[:n_obs] / counts[:n_obs].astype(float)
        n_exit = int(rb.n_exit.item()) if rb.n_exit is not None else 0
        rho_hat = fft_amp = None
        if record_fft:
            hat = torch.fft.fft(total[0], dim=-1)
            rho_hat = hat.cpu().numpy()
            fft_amp = hat.abs().cpu().numpy()
        var_list = None
        if record_var:
            var_list = var[0].cpu().numpy() if var is not None else np.zeros(M, dtype=float)
        self.last_run_info = dict(n_events=int(rb.n_events.item()), n_obs=n_obs, launches=launches,
                                  t_end=float(rb.t_end.item()), n_guard=int(rb.n_guard.item()),
                                  mode="philox" if native else "replay")
        return {
            "times_obs": times_obs,
            "pos_list": [out_pos[m, :counts[m]].astype(np.int64) if m < n_obs else None for m in range(M)],
            "rho_p_list": rho_p[0].cpu().numpy(),
            "rho_m_list": rho_m[0].cpu().numpy(),
            "total_list": total[0].cpu().numpy(),
            "particle_count_list": [int(counts[m]) if m < n_obs else None for m in range(M)],
            "bound_list": [(bounds[m, :counts[m]].copy() if bounds is not None else np.zeros(counts[m], dtype=bool))
                           if m < n_obs else None for m in range(M)],
            "m_local_list": rb.obs_m_local[0].cpu().numpy(),
            "m_global": m_global,
            "rho_hat_complex": rho_hat,
            "fft_amp_list": fft_amp,
            "var_list": var_list,
            "exit_times": rb.exit_t[0, :n_exit].cpu().numpy().tolist() if n_exit else [],
            "exit_positions": rb.exit_pos[0, :n_exit].cpu().numpy().astype(np.int64).tolist() if n_exit else [],
        }

    # per event the reference calls rng.exponential(1/R), rng.choice(n, p=...), rng.random() and, for a
    # diffusive hop only, rng.random() again (CLASS.py:358-362,378).  exponential(1.0) returns the standard
    # variate e with tau = (1/R)*e, and choice(n, p) consumes exactly one random().
    def _draw_triples(self, k):
        rng = self.rng
        out = np.empty(3 * k, dtype=np.float64)
        ex, rnd = rng.exponential, rng.random
        for i in range(k):
            out[3 * i] = ex(1.0)
            out[3 * i + 1] = rnd()
            out[3 * i + 2] = rnd()
        return out

    def _run_replay(self, rb):
        import torch

        rng = self.rng
        can_rewind = hasattr(rng, "bit_generator")
        chunk = (8192 if self.rate_diffusion == 0 else 512) if can_rewind else 1
        rb.pos_end.copy_(rb.pos0)
        rb.sigma_end.copy_(rb.sigma0)
        rb.pos0, rb.sigma0 = rb.pos_end, rb.sigma_end          # in-place state: each launch resumes
        rb.n_end.copy_(rb.n)
        rb.n = rb.n_end                                        # exits shrink n between launches
        rb.bound0 = rb.bound_end                               # None without anchors
        resume = dict(t_start=rb.t_end, obs_start=rb.n_obs, ev_start=rb.n_events)
        off = torch.tensor([0, 0], dtype=torch.int64, device=rb.dev)
        prefix = np.empty(0, dtype=np.float64)
        launches = 0
        while True:
            state = rng.bit_generator.state if can_rewind else None
            spec = self._draw_triples(chunk)
            draws = np.concatenate([prefix, spec])
            d = torch.from_numpy(draws).to(rb.dev)
            off[1] = draws.size
            rb.run_replay(d, off, resume=resume, spec_from=prefix.size)
            launches += 1
            status = int(rb.status.item())
            used = int(rb.draws_used.item())
            if status != capi.APS_RUN_DRAWS_EXHAUSTED:
                if can_rewind:                               # leave rng where the reference would
                    rng.bit_generator.state = state
                    self._draw_triples((used - prefix.size) // 3)
                return launches
            if used == draws.size:                           # ran out exactly at an event boundary
                prefix = np.empty(0, dtype=np.float64)
                continue
            # a diffusive event at offset `used` needs a 4th variate that the speculation did not draw
            k_done = (used - prefix.size) // 3
            if can_rewind:
                rng.bit_generator.state = state
                self._draw_triples(k_done)
                head = self._draw_triples(1)
            else:
                head = draws[used:used + 3]
            prefix = np.concatenate([head, [rng.random()]])

    # ---- plotting / animation: the reference's own post-processing of `out`, borrowed from its code ----
    def visualize_all(self, out, *args, **kwargs):
        """CLASS.py:561-661 (delegated to the reference's implementation, see reference_class())."""
        return reference_class().visualize_all(self, out, *args, **kwargs)

    def plot_individuals(self, out, *args, **kwargs):
        """CLASS.py:663-978 (delegated; returns the reference's `mean_v_eff`)."""
        return reference_class().plot_individuals(self, out, *args, **kwargs)

    def animate_profiles(self, out, *args, **kwargs):
        """CLASS.py:980-1093 (delegated)."""
        return reference_class().animate_profiles(self, out, *args, **kwargs)

    def step_gillespie(self, pos, sigma, bound, m_field, counts_p, counts_m, init_bin, exit_times,
                       exit_positions, exit_init_bin, t):
        """One event (CLASS.py:254-448) on the device.  Arrays are updated in place and returned like the
        reference does; `m_field` is honoured as given (it is an input of the reference's step).
        Systems with anchors (bind / unbind / exit events, `bound` flags) are supported by `run()` only."""
        n = sigma.size
        if n == 0:
            return pos, sigma, bound, np.inf, counts_p, counts_m
        if self._has_anchors:
            raise NotImplementedError("step_gillespie: the single-step entry does not carry anchor binding state; "
                                      "use run() for systems with anchor_positions (CLASS.py:307-348,418-436)")
        from .engine import ReplicaBatch
        import torch

        rb = ReplicaBatch(L=self.L, K=self.K, radius=self._radius, weights=self._weights, D=self.rate_diffusion,
                          lam=self.rate_active, T=np.inf, times_obs=[0.0], betas=[float(self.beta)], n=[n],
                          pos0=np.asarray(pos, np.int32).reshape(1, -1), sigma0=np.asarray(sigma, np.int8).reshape(1, -1),
                          record=0, crowding=self.crowding_suppresses_rates, dx=self.dx, periodic=self.periodic,
                          flip_tab=self._flip_tab)
        mf = torch.from_numpy(np.ascontiguousarray(m_field, dtype=np.float64)).to(rb.dev)
        one = torch.ones(1, dtype=torch.int32, device=rb.dev)
        draws = self._draw_triples(1)
        while True:
            d = torch.from_numpy(draws).to(rb.dev)
            off = torch.tensor([0, draws.size], dtype=torch.int64, device=rb.dev)
            rb.run_replay(d, off, max_events=1, resume=dict(obs_start=one, m_field_in=mf),
                          spec_from=0 if draws.size == 3 else -1)
            if int(rb.status.item()) == capi.APS_RUN_DRAWS_EXHAUSTED and draws.size == 3:
                draws = np.concatenate([draws, [self.rng.random()]])
                continue
            break
        if int(rb.status.item()) == capi.APS_RUN_EMPTY:
            return pos, sigma, bound, np.inf, counts_p, counts_m
        tau = float(rb.t_end.item())
        new_pos = rb.pos_end[0].cpu().numpy()
        new_sig = rb.sigma_end[0].cpu().numpy()
        i = int(np.nonzero((new_pos != pos) | (new_sig != sigma))[0][0])
        old = int(pos[i])
        cnt_old = counts_p if sigma[i] == 1 else counts_m
        if new_sig[i] != sigma[i]:
            cnt_old[old] -= 1
            (counts_m if sigma[i] == 1 else counts_p)[old] += 1
        else:
            cnt_old[old] -= 1
            cnt_old[int(new_pos[i])] += 1
        pos[i] = new_pos[i]
        sigma[i] = new_sig[i]
        return pos, sigma, bound, tau, counts_p, counts_m, exit_times, exit_positions, exit_init_bin
