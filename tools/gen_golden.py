#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by executing the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference has no tests or
golden vectors of its own, so parity is pinned on what its code does here:
`ParticleSystem` is imported from /root/reference with matplotlib/vispy stubbed (they are
plot-only imports, PARTICLE_solver_CLASS.py:4-7) and driven through its public API with a
recording wrapper around a seeded numpy Generator passed via the constructor's `rng=` seam
(PARTICLE_solver_CLASS.py:26,75-78).

For each case we store: constructor/run parameters, Gaussian taps (scipy's), initial state,
the variate log (standard exponential + uniforms in call order), the per-event trace
(particle, kind, new site) and the returned `out` arrays.  Every case is also run a second
time with a plain, unwrapped `np.random.default_rng(seed)` to prove that the wrapper does not
change the reference's behaviour.

Also stores outputs of the sweep drivers' per-run reducers (extracted by AST from the driver
file, because the drivers execute their sweep at import time).

Usage: python tools/gen_golden.py [case ...]
"""
from __future__ import annotations

import ast
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "vispy", "vispy.app", "vispy.scene", "vispy.io"]:
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from PARTICLE_solver_CLASS import ParticleSystem  # type: ignore

    return ParticleSystem


def extract_functions(path, names):
    """exec selected top-level function definitions of a driver script (no sweep is run)."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    return ns


class RecordingRNG:
    """Duck-typed rng for ParticleSystem: forwards to a numpy Generator and logs the variates of
    the three per-event calls (exponential, choice(n, p=...), random)."""

    def __init__(self, gen):
        self.gen = gen
        self.log = []

    def exponential(self, scale=1.0):
        e = self.gen.standard_exponential()
        self.log.append(e)
        return scale * e

    def choice(self, a, size=None, replace=True, p=None):
        if p is not None and size is None:
            u = self.gen.random()
            self.log.append(u)
            cdf = np.asarray(p).cumsum()
            cdf /= cdf[-1]
            self.last_choice = int(cdf.searchsorted(u, side="right"))
            return self.last_choice
        return self.gen.choice(a, size=size, replace=replace, p=p)

    def random(self):
        u = self.gen.random()
        self.log.append(u)
        return u

    def poisson(self, lam):
        return self.gen.poisson(lam)


def exp_gradient(L, N, frac_plus, decay_length):
    xs = np.arange(L) / float(L)
    plus = np.exp(-xs / decay_length)
    minus = 0.05 * np.ones_like(xs)
    rho_plus = N * frac_plus * (plus / plus.sum())
    rho_minus = N * (1 - frac_plus) * (minus / minus.sum())
    return rho_plus, rho_minus


def profile_callables(rho_plus, rho_minus):
    L = len(rho_plus)

    def fp(x):
        return float(rho_plus[int(np.clip(np.round(x * L), 0, L - 1))])

    def fm(x):
        return float(rho_minus[int(np.clip(np.round(x * L), 0, L - 1))])

    return fp, fm


BASE = dict(flip_rate_fn=None, minus_anchor=True, periodic=False, immobilize_when_anchored=True,
            anchor_radius=0.003, anchor_positions=None, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)


def cases():
    c = {}
    c["c1_exclusion"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0, rate_active=5, beta=0.7, init="fixed", N=750,
                                     scale_rates=False, local_kernel_sigma=0.002, site_capacity=3),
                             run=dict(T=1.0, obs_dt=0.25, record_fft=True, record_var=True), seed=101)
    sw = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, init="poisson", N=500, scale_rates=False,
              local_kernel_sigma=0.005, site_capacity=1)
    for b, name in [(0.0, "c2_sweep_b0"), (3.0, "c2_sweep_b3")]:
        c[name] = dict(ps=dict(sw, beta=b), profile=dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.35),
                       run=dict(T=1.5, obs_dt=0.1, record_fft=True, record_var=True), seed=202 + int(b))
    c["c3_local"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.05, rate_active=5, beta=1.5, init="fixed", N=900,
                                 scale_rates=False, local_kernel_sigma=0.005, site_capacity=1),
                         run=dict(T=2.0, obs_dt=0.5, record_fft=True, record_var=True), seed=303)
    c["c4_double"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.005, rate_active=10, beta=1.5, init="poisson", N=500,
                                  scale_rates=False, local_kernel_sigma=0.02, site_capacity=1),
                          profile=dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.2),
                          run=dict(T=1.0, obs_dt=0.1, record_fft=False, record_var=False), seed=404)
    c["global_sigma0"] = dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.002, rate_active=5, beta=2.0, init="poisson", N=500,
                                      scale_rates=False, local_kernel_sigma=0.0, site_capacity=1),
                              profile=dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.35),
                              run=dict(T=1.5, obs_dt=0.1, record_fft=False, record_var=True), seed=505)
    c["huge_sigma"] = dict(ps=dict(L=100, xlim=1, rate_diffusion=0.5, rate_active=2, beta=1.0, init="fixed", N=40,
                                   scale_rates=False, local_kernel_sigma=0.3, site_capacity=1),
                           run=dict(T=2.0, obs_dt=0.25, record_fft=False, record_var=False), seed=606)
    c["tiny_diffusive"] = dict(ps=dict(L=16, xlim=1, rate_diffusion=3.0, rate_active=1, beta=0.5, init="fixed", N=8,
                                       scale_rates=False, local_kernel_sigma=0.1, site_capacity=2),
                               run=dict(T=3.0, obs_dt=0.2, record_fft=True, record_var=True), seed=707)
    c["k1_dense"] = dict(ps=dict(L=64, xlim=1, rate_diffusion=1.0, rate_active=3, beta=2.5, init="fixed", N=60,
                                 scale_rates=False, local_kernel_sigma=0.03, site_capacity=1),
                         run=dict(T=3.0, obs_dt=0.25, record_fft=False, record_var=False), seed=808)
    c["r0_local"] = dict(ps=dict(L=50, xlim=1, rate_diffusion=0.7, rate_active=2, beta=1.2, init="fixed", N=30,
                                 scale_rates=False, local_kernel_sigma=0.001, site_capacity=2),
                         run=dict(T=3.0, obs_dt=0.5, record_fft=False, record_var=False), seed=909)
    c["crowding"] = dict(ps=dict(L=64, xlim=1, rate_diffusion=0.8, rate_active=3, beta=1.0, init="fixed", N=100,
                                 scale_rates=False, local_kernel_sigma=0.04, site_capacity=3,
                                 crowding_suppresses_rates=True),
                         run=dict(T=2.0, obs_dt=0.25, record_fft=False, record_var=False), seed=1010)
    c["scale_rates"] = dict(ps=dict(L=32, xlim=1, rate_diffusion=0.0005, rate_active=0.05, beta=0.8, init="fixed", N=20,
                                    scale_rates=True, local_kernel_sigma=0.05, site_capacity=1),
                            run=dict(T=4.0, obs_dt=0.5, record_fft=False, record_var=False), seed=1111)
    c["multi_obs_per_event"] = dict(ps=dict(L=16, xlim=1, rate_diffusion=0.05, rate_active=0.1, beta=0.3, init="fixed", N=3,
                                            scale_rates=False, local_kernel_sigma=0.1, site_capacity=1),
                                    run=dict(T=5.0, obs_dt=0.01, record_fft=False, record_var=False), seed=1212)
    for k, seed in enumerate([1313, 1314, 1315, 1316]):
        c[f"sparse_events_{k}"] = dict(ps=dict(L=8, xlim=1, rate_diffusion=0.1, rate_active=0.1, beta=0.0, init="fixed", N=2,
                                               scale_rates=False, local_kernel_sigma=0.2, site_capacity=1),
                                       run=dict(T=1.0, obs_dt=0.9, record_fft=False, record_var=False), seed=seed)
    anch = dict(anchor_positions=[0.25, 0.60, 0.80], anchor_radius=0.01, k_on=10, k_off=5, k_exit=5)
    c["anchors_k3"] = dict(ps=dict(L=200, xlim=1, rate_diffusion=0.5, rate_active=3, beta=1.0, init="fixed", N=150,
                                   scale_rates=False, local_kernel_sigma=0.02, site_capacity=3, **anch),
                           run=dict(T=2.0, obs_dt=0.25, record_fft=False, record_var=False), seed=1515)
    c["anchors_free_flip"] = dict(ps=dict(L=120, xlim=1, rate_diffusion=0.3, rate_active=2, beta=0.5, init="fixed", N=90,
                                          scale_rates=False, local_kernel_sigma=0.03, site_capacity=2,
                                          immobilize_when_anchored=False, suppress_flip_when_bound=False, **anch),
                                  run=dict(T=2.0, obs_dt=0.25, record_fft=False, record_var=False), seed=1616)
    c["anchors_crowding_global"] = dict(ps=dict(L=100, xlim=1, rate_diffusion=0.4, rate_active=2, beta=1.5, init="fixed", N=120,
                                                scale_rates=False, local_kernel_sigma=0.0, site_capacity=3,
                                                crowding_suppresses_rates=True, **dict(anch, k_exit=20)),
                                        run=dict(T=3.0, obs_dt=0.25, record_fft=False, record_var=False), seed=1717)
    # periodic=True (ring): the reference applies the kernel by FFT (CLASS.py:224-227); hops wrap (:278-288)
    c["periodic_k1"] = dict(ps=dict(L=200, xlim=1, rate_diffusion=0.4, rate_active=3, beta=1.5, init="fixed", N=120,
                                    scale_rates=False, local_kernel_sigma=0.02, site_capacity=1, periodic=True),
                            run=dict(T=2.0, obs_dt=0.25, record_fft=False, record_var=False), seed=1818)
    c["periodic_k2_crowding"] = dict(ps=dict(L=64, xlim=1, rate_diffusion=0.8, rate_active=2, beta=0.8, init="fixed", N=70,
                                             scale_rates=False, local_kernel_sigma=0.03, site_capacity=2, periodic=True,
                                             crowding_suppresses_rates=True),
                                     run=dict(T=2.0, obs_dt=0.25, record_fft=False, record_var=False), seed=1919)
    c["periodic_global"] = dict(ps=dict(L=50, xlim=1, rate_diffusion=0.6, rate_active=2, beta=1.2, init="fixed", N=30,
                                        scale_rates=False, local_kernel_sigma=0.0, site_capacity=1, periodic=True),
                                run=dict(T=3.0, obs_dt=0.5, record_fft=False, record_var=False), seed=2020)
    # custom flip_rate_fn (CLASS.py:59-62): named callables of tests/common.py FLIP_FNS
    c["custom_flip_local"] = dict(ps=dict(L=200, xlim=1, rate_diffusion=0.3, rate_active=3, beta=1.5, init="fixed", N=120,
                                          scale_rates=False, local_kernel_sigma=0.02, site_capacity=1),
                                  flip="glauber_1p5", run=dict(T=3.0, obs_dt=0.25, record_fft=False, record_var=False), seed=2121)
    c["custom_flip_global_k2"] = dict(ps=dict(L=100, xlim=1, rate_diffusion=0.4, rate_active=2, beta=0.8, init="fixed", N=110,
                                              scale_rates=False, local_kernel_sigma=0.0, site_capacity=2),
                                      flip="exp_quadratic", run=dict(T=3.0, obs_dt=0.25, record_fft=False, record_var=False), seed=2222)
    c["poisson_k2"] = dict(ps=dict(L=200, xlim=1, rate_diffusion=0.3, rate_active=4, beta=1.8, init="poisson", N=260,
                                   scale_rates=False, local_kernel_sigma=0.01, site_capacity=2),
                           profile=dict(L=200, N=260, frac_plus=0.6, decay_plus=0.5),
                           run=dict(T=1.0, obs_dt=0.125, record_fft=True, record_var=True), seed=1414)
    return c


def build_ps(PS, spec, rng):
    kw = dict(BASE)
    kw.update(spec["ps"])
    if spec.get("flip"):
        sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
        from common import FLIP_FNS
        kw["flip_rate_fn"] = FLIP_FNS[spec["flip"]]
    if "profile" in spec:
        p = spec["profile"]
        rp, rm = exp_gradient(p["L"], p["N"], p["frac_plus"], p["decay_plus"])
        fp, fm = profile_callables(rp, rm)
        kw.update(rho0_plus=fp, rho0_minus=fm)
    return PS(rng=rng, **kw)


def run_case(PS, name, spec):
    from scipy.ndimage import _filters  # scipy's own tap generator

    seed = spec["seed"]
    # (1) plain reference run, unwrapped numpy Generator
    ps_plain = build_ps(PS, spec, np.random.default_rng(seed))
    out_plain = ps_plain.run(**spec["run"])

    # (2) recorded run: capture initial state, draws and per-event trace
    rec = RecordingRNG(np.random.default_rng(seed))
    ps = build_ps(PS, spec, rec)
    captured = {}
    orig_init = ps.init_particles

    def init_hook():
        pos, sigma = orig_init()
        captured["pos0"] = pos.copy()
        captured["sigma0"] = sigma.copy()
        return pos, sigma

    ps.init_particles = init_hook
    trace = []
    orig_step = ps.step_gillespie

    def step_hook(pos, sigma, bound, *a):
        p0, s0, b0, nlog = pos.copy(), sigma.copy(), bound.copy(), len(rec.log)
        res = orig_step(pos, sigma, bound, *a)
        used = len(rec.log) - nlog
        i = rec.last_choice                       # the particle rng.choice picked (CLASS.py:360)
        if res[0].size < p0.size:
            assert np.array_equal(res[0], np.delete(p0, i)) and np.array_equal(res[1], np.delete(s0, i))
            trace.append((i, 6, int(p0[i])))
        elif res[2][i] != b0[i]:
            trace.append((i, 4 if res[2][i] else 5, -1))
        elif res[1][i] != s0[i]:
            trace.append((i, 3, -1))
        elif res[0][i] != p0[i]:
            step = int(res[0][i] - p0[i])
            kind = (0 if step < 0 else 1) if used == 4 else 2
            trace.append((i, kind, int(res[0][i])))
        else:
            raise RuntimeError("event did not change the selected particle")
        assert (res[0] != p0[:res[0].size]).sum() <= (res[0].size if res[0].size < p0.size else 1)
        return res

    ps.step_gillespie = step_hook
    out = ps.run(**spec["run"])

    # the wrapper must be behaviour-neutral
    for k in ["rho_p_list", "rho_m_list", "total_list", "m_local_list", "m_global"]:
        assert np.array_equal(out[k], out_plain[k]), (name, k)
    n_obs = sum(p is not None for p in out["pos_list"])
    for a, b in zip(out["pos_list"][:n_obs], out_plain["pos_list"][:n_obs]):
        assert np.array_equal(a, b)

    n = captured["pos0"].size
    pos_obs = np.full((len(out["pos_list"]), n), -1, dtype=np.int32)
    bound_obs = np.zeros((len(out["pos_list"]), n), dtype=np.int8)
    count_obs = np.zeros(len(out["pos_list"]), dtype=np.int32)
    for m in range(n_obs):
        k = out["pos_list"][m].size
        pos_obs[m, :k] = out["pos_list"][m]
        bound_obs[m, :k] = out["bound_list"][m]
        count_obs[m] = out["particle_count_list"][m]
    radius = -1
    weights = np.zeros(0)
    if ps.local_kernel_sigma > 0 and ps.periodic:
        # the reference's own ring kernel (ps._kernel, CLASS.py:111-121), truncated where the discarded mass <= 1e-22
        kern = ps._kernel
        radius = next(r for r in range((ps.L - 1) // 2 + 1) if kern[r + 1:ps.L - r].sum() <= 1e-22)
        weights = np.concatenate([kern[radius:0:-1], kern[:radius + 1]]).copy()
    elif ps.local_kernel_sigma > 0:
        sd = float(ps._sigma_grid)
        radius = int(4.0 * sd + 0.5)
        weights = _filters._gaussian_kernel1d(sd, 0, radius)[::-1].copy()
    meta = dict(name=name, ps=spec["ps"], run=spec["run"], seed=seed, profile=spec.get("profile"), flip=spec.get("flip"),
                dx=ps.dx, rate_diffusion=ps.rate_diffusion, rate_active=ps.rate_active, K=ps.K, L=ps.L,
                radius=radius, n=int(n), n_obs=int(n_obs), n_events=len(trace),
                periodic=bool(ps.periodic), anchors=bool(ps.is_anchor_site.any()), k_on=float(ps.k_on), k_off=float(ps.k_off), k_exit=float(ps.k_exit),
                suppress=bool(ps.suppress_flip_when_bound), immobilize=bool(ps.immobilize_when_anchored),
                numpy=np.__version__, scipy=__import__("scipy").__version__)
    save = dict(
        meta=np.array(json.dumps(meta)),
        weights=weights,
        times_obs=out["times_obs"],
        pos0=captured["pos0"].astype(np.int32),
        sigma0=captured["sigma0"].astype(np.int8),
        draws=np.asarray(rec.log, dtype=np.float64),
        trace=np.asarray(trace, dtype=np.int32).reshape(-1, 3),
        rho_p_list=out["rho_p_list"], rho_m_list=out["rho_m_list"], total_list=out["total_list"],
        m_local_list=out["m_local_list"], m_global=out["m_global"], pos_obs=pos_obs,
        bound_obs=bound_obs, count_obs=count_obs, anchor_mask=ps.is_anchor_site.astype(np.uint8),
        exit_times=np.asarray(out["exit_times"], dtype=np.float64), exit_positions=np.asarray(out["exit_positions"], dtype=np.int32),
    )
    if out["var_list"] is not None:
        save["var_list"] = out["var_list"]
    if out["fft_amp_list"] is not None:
        save["fft_amp_head"] = out["fft_amp_list"][:, :32].copy()
        save["rho_hat_head"] = out["rho_hat_complex"][:, :32].copy()
    if "profile" in spec:
        p = spec["profile"]
        rp, rm = exp_gradient(p["L"], p["N"], p["frac_plus"], p["decay_plus"])
        save["rho0_plus"], save["rho0_minus"] = rp, rm
        assert np.array_equal(ps.rho0_plus, rp) and np.array_equal(ps.rho0_minus, rm)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **save)
    print(f"{name}: n={n} events={len(trace)} draws={len(rec.log)} n_obs={n_obs}/{len(out['pos_list'])}")
    return ps, out


def cases_reducers():
    return {tag: dict(ps=dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, beta=beta, init="poisson", N=500,
                              scale_rates=False, local_kernel_sigma=0.005, site_capacity=1),
                      profile=dict(L=1000, N=500, frac_plus=0.75, decay_plus=0.35),
                      run=dict(T=3.0, obs_dt=0.1, record_fft=True, record_var=True), seed=seed)
            for tag, beta, seed in [("b0", 0.5, 77), ("b2", 2.0, 78)]}


def reducer_golden(PS):
    """Golden outputs of the sweep driver's per-run reducers on a short sweep_beta-like run."""
    names = ["compute_v_eff_and_window", "compute_rho_eff", "compute_blocking_probability",
             "compute_mean_magnetizatoin", "compute_D_eff_active", "make_exp_gradient"]
    ns = extract_functions(os.path.join(REF, "PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py"), names)
    # our exp_gradient must match the driver's make_exp_gradient
    r = ns["make_exp_gradient"](L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
    rp, rm = exp_gradient(1000, 500, 0.75, 0.35)
    assert np.array_equal(r[2], rp) and np.array_equal(r[3], rm)
    assert all(r[0](i / 1000) == rp[i] and r[1](i / 1000) == rm[i] for i in range(1000))
    res = {}
    for tag, spec in cases_reducers().items():
        ps, out = run_case(PS, f"reducers_{tag}", spec)
        mean_v, v_eff, times, si, ei, frac_b = ns["compute_v_eff_and_window"](
            out, ps, boundary_xmin=0.99, max_buondary_fraction=0.06, min_window_fraction=0.10)
        res[tag] = dict(mean_v=mean_v, si=int(si), ei=int(ei),
                        D_eff=float(ns["compute_D_eff_active"](out, ps, start_idx=si, end_idx=ei)),
                        m_mean=ns["compute_mean_magnetizatoin"](out, si, ei),
                        rho_eff=ns["compute_rho_eff"](out, si, ei),
                        block=float(ns["compute_blocking_probability"](out, si, ei)),
                        v_eff=v_eff.tolist(), frac_boundary=frac_b.tolist())
    json.dump(res, open(os.path.join(OUT, "reducers.json"), "w"), indent=1)
    print("reducers:", {k: {kk: vv for kk, vv in v.items() if not isinstance(vv, list)} for k, v in res.items()})


def structure_golden(PS):
    """Golden outputs of the local_structure driver's per-run analyses (local_structure.py:55-103,195-265) on the two
    recorded reducer runs (record_fft=True), executed from the reference's own function definitions."""
    names = ["extract_structure_observables_from_out", "time_to_pattern", "ensemble_time_to_pattern",
             "cluster_size_distribution", "temporal_autocorrelation", "lowk_variance_time", "spectral_entropy",
             "mode_competition_ratio", "extract_growth_rate"]
    ns = extract_functions(os.path.join(REF, "PARTICLE_solver_BIOLOGY_local_structure.py"), names)
    cases = cases_reducers()
    res, outs = {}, []
    for tag in ["b0", "b2"]:
        ps, out = run_case(PS, f"reducers_{tag}", cases[tag])
        outs.append(out)
        obs = ns["extract_structure_observables_from_out"](out, start_fraction=0.5, k_max=None)
        thr_k = {k: float(np.sort(out["fft_amp_list"][:, k])[-3:-1].mean()) for k in (1, 2, 3, 5)}   # exceeded by two rows only
        thr = thr_k[2]
        d = {k: (float(v) if np.isscalar(v) or np.ndim(v) == 0 else None) for k, v in obs.items()}
        d.update(fft_mean_head=obs["fft_mean"][:40].tolist(), fft_std_head=obs["fft_std"][:40].tolist(),
                 dominant_k=int(obs["dominant_k"]), threshold=thr,
                 thresholds={str(k): v for k, v in thr_k.items()},
                 time_to_pattern={str(k): float(ns["time_to_pattern"](out, threshold=v, k=k)) for k, v in thr_k.items()},
                 time_to_pattern_never=float(ns["time_to_pattern"](out, threshold=1e9, k=2)),
                 lowk_variance_time=ns["lowk_variance_time"](out, k_cut=25).tolist(),
                 autocorr_lag1=float(ns["temporal_autocorrelation"](out, lag=1)),
                 autocorr_lag3=float(ns["temporal_autocorrelation"](out, lag=3)),
                 growth_rate=float(ns["extract_growth_rate"](out, k=1, t_min=0.5, t_max=2.5, amp_min=1e-4)),
                 growth_rate_k3=float(ns["extract_growth_rate"](out, k=3, t_min=0.0, t_max=None, amp_min=1e-4)),
                 growth_rate_nan=float(ns["extract_growth_rate"](out, k=1, t_min=2.85, t_max=None, amp_min=1e-4)),
                 spectral_entropy=float(ns["spectral_entropy"](obs["fft_mean"], k_max=25)),
                 spectral_entropy_full=float(ns["spectral_entropy"](obs["fft_mean"])),
                 mode_competition=float(ns["mode_competition_ratio"](obs["fft_mean"])),
                 clusters=ns["cluster_size_distribution"](out["total_list"][-1], 0.5 * out["total_list"][-1].max()).tolist())
        res[tag] = d
    thr = min(res["b0"]["threshold"], res["b2"]["threshold"])
    mean_t, se_t = ns["ensemble_time_to_pattern"](outs, k=2, threshold=thr)
    res["ensemble"] = dict(threshold=thr, mean=float(mean_t), se=float(se_t))
    json.dump(res, open(os.path.join(OUT, "structure.json"), "w"), indent=1)
    print("structure:", {k: {kk: vv for kk, vv in v.items() if not isinstance(vv, list)} for k, v in res.items()})


STAT_PS = dict(L=100, xlim=1, rate_diffusion=0.3, rate_active=3, init="fixed", N=40, scale_rates=False,
               local_kernel_sigma=0.03, site_capacity=1)
STAT_RUN = dict(T=4.0, obs_dt=0.25, record_fft=False, record_var=False)


def stat_fixture(PS, n_runs=300):
    """Ensemble statistics of the unmodified reference (seeded numpy Generators) for the native-mode
    statistical parity test: per-site mean/variance over replicas of the time-averaged (second half)
    rho_plus / rho_minus profiles and the per-replica time-averaged m_global."""
    save = dict(meta=np.array(json.dumps(dict(ps=STAT_PS, run=STAT_RUN, n_runs=n_runs, betas=[0.5, 2.0]))))
    for bi, beta in enumerate([0.5, 2.0]):
        prof_p, prof_m, mbar, nev = [], [], [], []
        for r in range(n_runs):
            kw = dict(BASE); kw.update(STAT_PS)
            rec = RecordingRNG(np.random.default_rng(900000 + 1000 * bi + r))
            ps = PS(beta=beta, rng=rec, **kw)
            out = ps.run(**STAT_RUN)
            M = len(out["times_obs"])
            assert all(p is not None for p in out["pos_list"])
            prof_p.append(out["rho_p_list"][M // 2:].mean(0)); prof_m.append(out["rho_m_list"][M // 2:].mean(0))
            mbar.append(out["m_global"][M // 2:].mean())
            nev.append(sum(1 for _ in rec.log) )
        prof_p, prof_m = np.array(prof_p), np.array(prof_m)
        save[f"b{bi}_rho_p_mean"] = prof_p.mean(0); save[f"b{bi}_rho_p_var"] = prof_p.var(0, ddof=1)
        save[f"b{bi}_rho_m_mean"] = prof_m.mean(0); save[f"b{bi}_rho_m_var"] = prof_m.var(0, ddof=1)
        save[f"b{bi}_mbar"] = np.array(mbar)
        print(f"stat beta={beta}: mean m = {np.mean(mbar):.4f} +- {np.std(mbar) / np.sqrt(n_runs):.4f}")
    np.savez_compressed(os.path.join(OUT, "stat_ensemble.npz"), **save)


PDE_CASE = dict(L_lattice=131072, peak_density=0.02, frac_plus=0.8, D=0.5, lam=20.0, beta=1.5, T=1760.0,
                init="exp(-|x-0.5|/0.05) bump (IMEXPDE.initialize mode='poisson', noise=0)")


def pde_fixture():
    """BASELINE config 5 comparison target: the reference's IMEXPDE (IMEX_PDE_solver_class.py, unmodified,
    matplotlib stubbed) run on the hydrodynamic scaling of the K2 lattice of PDE_CASE:
        x = site / L_lattice,  gamma = D / L_lattice^2,  lam = lambda / L_lattice,  same time unit,
    in the setting where the PDE and the particle model describe the same dynamics: only '+'
    particles are advected (active_model='anchored_minus'), reflecting walls (bc='neumann'), global
    magnetisation (kernel_sigma = 1e5: flat kernel, as IMEX_PDE_solver_run_sweep.py uses), dilute lattice (the PDE has no exclusion term)."""
    for m in ["matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from IMEX_PDE_solver_class import IMEXPDE  # type: ignore
    c = PDE_CASE
    Lk = c["L_lattice"]
    pde = IMEXPDE(L=1000, xlim=1.0, T=c["T"], dt=0.1, gamma=c["D"] / Lk ** 2, lam=c["lam"] / Lk, beta=c["beta"],
                  bc="neumann", active_model="anchored_minus", gaussian_kernel=True, kernel_sigma=1e5,
                  snapshot_interval=10 ** 9, outdir="/tmp/imex_fixture", seed=1)
    pde.initialize(mode="poisson", rho0=1.0, noise=0.0, n_tracers=4)
    tot = pde.rho_p + pde.rho_m
    pde.rho_p = tot * c["frac_plus"]
    pde.rho_m = tot * (1.0 - c["frac_plus"])
    with np.errstate(all="ignore"):
        pde.solve()
    out = pde.get_output()
    np.savez_compressed(os.path.join(OUT, "stat_pde_config5.npz"), meta=np.array(json.dumps(c)),
                        rho_p=out["rho_p"], rho_m=out["rho_m"], m_series=out["m_series"][::100])
    tot = out["rho_p"] + out["rho_m"]
    x = (np.arange(1000) + 0.0) / 1000
    print("pde: mass", tot.sum(), "m_final", out["m_series"][-1], "mean x", (tot * x).sum() / tot.sum(),
          "std x", np.sqrt((tot * x * x).sum() / tot.sum() - ((tot * x).sum() / tot.sum()) ** 2))


PDE_STEP_CASES = {
    "pde_periodic_kernel": dict(ctor=dict(L=200, xlim=1.0, T=0.15, dt=5e-4, gamma=2e-3, lam=0.6, beta=2.0, bc="periodic",
                                          active_model="bidirectional", gaussian_kernel=True, kernel_sigma=0.02,
                                          snapshot_interval=50, seed=5),
                                init=dict(mode="homogeneous", rho0=1.0, noise=0.3, n_tracers=64)),
    "pde_full_ring_kernel": dict(ctor=dict(L=128, xlim=1.0, T=0.1, dt=5e-4, gamma=0.0, lam=0.6, beta=1.0, bc="periodic",
                                           active_model="bidirectional", gaussian_kernel=True, kernel_sigma=1e5 - 10,
                                           snapshot_interval=40, seed=6),
                                 init=dict(mode="homogeneous", rho0=1.0, noise=0.3, n_tracers=32)),
    "pde_neumann_pointwise": dict(ctor=dict(L=100, xlim=1.0, T=0.15, dt=5e-4, gamma=0.2, lam=0.6, beta=1.5, bc="neumann",
                                            active_model="bidirectional", gaussian_kernel=False, kernel_sigma=0.02,
                                            snapshot_interval=100, seed=7),
                                  init=dict(mode="poisson", rho0=1.0, noise=0.05, n_tracers=16)),
    "pde_anchored_minus": dict(ctor=dict(L=150, xlim=1.0, T=0.2, dt=1e-3, gamma=0.01, lam=0.8, beta=2.5, bc="periodic",
                                         active_model="anchored_minus", gaussian_kernel=True, kernel_sigma=0.05,
                                         snapshot_interval=25, seed=8),
                               init=dict(mode="homogeneous", rho0=1.0, noise=0.2, n_tracers=16)),
    "pde_neumann_anchored": dict(ctor=dict(L=96, xlim=1.0, T=0.1, dt=5e-4, gamma=0.05, lam=1.0, beta=0.7, bc="neumann",
                                           active_model="anchored_minus", gaussian_kernel=True, kernel_sigma=0.03,
                                           snapshot_interval=10, seed=9),
                                 init=dict(mode="poisson", rho0=1.0, noise=0.0, n_tracers=8)),
}


def pde_step_fixtures():
    """Short runs of the unmodified IMEXPDE (IMEX_PDE_solver_class.py; matplotlib stubbed) through its public API:
    initial state as initialize() leaves it, final fields, per-step diagnostics and snapshots (the deterministic part of
    solve(); tracer outputs depend on numpy's global stream inside the time loop and are compared statistically)."""
    for m in ["matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from IMEX_PDE_solver_class import IMEXPDE  # type: ignore
    for name, spec in PDE_STEP_CASES.items():
        pde = IMEXPDE(outdir="/tmp/imex_fixture", **spec["ctor"])
        pde.initialize(**spec["init"])
        init = dict(rho_p0=pde.rho_p.copy(), rho_m0=pde.rho_m.copy(), tracers0=pde.tracers_unwrapped.copy(),
                    tracer_state0=pde.tracer_state.astype(np.int8))
        with np.errstate(all="ignore"):
            pde.solve()
        out = pde.get_output()
        meta = dict(ctor={k: v for k, v in spec["ctor"].items()}, init=spec["init"], nsteps=int(pde.nsteps), numpy=np.__version__)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), meta=np.array(json.dumps(meta)), **init,
                            rho_p=out["rho_p"], rho_m=out["rho_m"], m_series=out["m_series"], var_series=out["var_series"],
                            snapshots=out["snapshots"], m_snapshots=out["m_snapshots"], times=out["times"],
                            v_eff_series=out["v_eff_series"], D_eff_series=out["D_eff_series"],
                            fft_amp_head=out["fft_amp"][:: max(1, pde.nsteps // 10), :16].copy())
        print(f"{name}: nsteps={pde.nsteps} mass={float((out['rho_p'] + out['rho_m']).sum()):.12f} m_end={out['m_series'][-1]:.6f}")


def main():
    os.makedirs(OUT, exist_ok=True)
    PS = import_reference()
    want = sys.argv[1:]
    for name, spec in cases().items():
        if want and name not in want:
            continue
        run_case(PS, name, spec)
    if not want or "reducers" in want:
        reducer_golden(PS)
    if not want or "structure" in want:
        structure_golden(PS)
    if not want or "stat" in want:
        stat_fixture(PS)
    if not want or "pde" in want:
        pde_fixture()
    if not want or "pde_step" in want:
        pde_step_fixtures()


if __name__ == "__main__":
    main()
