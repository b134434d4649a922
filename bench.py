#!/usr/bin/env python
"""Benchmark contract (see the task statement).

Metric (BASELINE.json): particle update attempts / second over a beta-sweep ensemble.  One exact-Gillespie event ==
one particle update attempt (every event is an accepted update, CLASS.py:351-367); for the sublattice kernel (K2) one
attempt == one particle visited in one pass.

Workloads (`--workload`, default config2; one contract JSON line per invocation):
  config2  BASELINE config 2 — PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta parameters (sweep_beta.py:829-878), 64 beta x 64
           replicas = 4096 independent ParticleSystem.run(T=20, obs_dt=0.1) per GPU (weak scaling).  The line also carries the
           K2 roofline measured at config 5's operating point and, at N > 1, the strong-scaling figure of the 4096-replica sweep.
  config3  BASELINE config 3 — PARTICLE_solver_BIOLOGY_local_structure parameters (:675-726): N=900 'fixed', T=40, obs_dt=1,
           64 beta x 64 replicas per GPU (weak), m_local rows recorded, structure analyses (cuFFT) on the device.
  config4  BASELINE config 4 — (density, beta) grid of PARTICLE_solver_BIOLOGY_EXCLUSION_double_sweep (:666-715): 16 x 16 points
           x 128 replicas = 32 768 replicas IN TOTAL at every N (strong scaling), sigma=0.02 (r=80), T=10.
  k2       BASELINE config 5 — one lattice of 2^26 sites IN TOTAL (strong scaling), local field sigma = 5 sites, dt = 0.005,
           slab decomposition with the in-kernel NVLink exchange; also L = 2^30 and the global field in `config`.

A step = one pass of the hot path over the whole workload.  `value` times it with inputs resident in HBM (CUDA events, max over
ranks); `e2e` times the public call (`launcher.sweep_over_betas`, `sweep_betas_for_structures`, `double_sweep`,
`SublatticeLattice`) from host parameters / host buffers to host results (wall clock, H2D and D2H inside the timed region).

`--impl reference` times the reference's own CPU implementation on the host cores, rank 0 only: for the ensemble workloads the
UNMODIFIED numpy `ParticleSystem.run` from baseline/_ref (tools/install_reference.py) in one process per core (kind
"reference"), else — no install on this box — the C restatement of the reference algorithm (oracle/, kind "port").  K2 has no
reference counterpart (the reference cannot run lattices beyond shared-memory size): its CPU arm is the oracle's sequential
restatement of the same update rule (kind "port").
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "particle_update_attempts_per_sec"
UNIT = "events/s"
N_BETA, REPS_PER_BETA = 64, 64
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

COMMON = dict(flip_rate_fn=None, minus_anchor=True, periodic=False, immobilize_when_anchored=True, anchor_radius=0.003,
              anchor_positions=None, crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0, scale_rates=False, L=1000, xlim=1,
              site_capacity=1)
WORKLOADS = {
    "config2": dict(
        ps=dict(COMMON, rate_diffusion=0.02, rate_active=5, init="poisson", N=500, local_kernel_sigma=0.005),   # sweep_beta.py:837-857
        run=dict(T=20, obs_dt=0.1, record_fft=True, record_var=True),                                          # :829-834
        profile=dict(N=500, frac_plus=0.75, decay_plus=0.35, decay_minus=0.2),                                 # :859-878
        scaling="weak",
        text="BASELINE config 2: sweep_beta ensemble, 64 beta x 64 replicas per GPU, L=1000, Poisson init N=500, K=1, "
             "sigma=0.005 (r=20), D=0.02, lambda=5, run(T=20, obs_dt=0.1)"),
    "config3": dict(
        ps=dict(COMMON, rate_diffusion=0.05, rate_active=5, init="fixed", N=900, local_kernel_sigma=0.005),    # local_structure.py:686-706
        run=dict(T=40, obs_dt=1, record_fft=True, record_var=True),                                            # :675-684
        profile=None, scaling="weak",
        text="BASELINE config 3: local_structure ensemble, 64 beta x 64 replicas per GPU, L=1000, N=900 'fixed', K=1, "
             "sigma=0.005 (r=20), D=0.05, lambda=5, run(T=40, obs_dt=1), m_local rows + structure observables (FFT) on the device"),
    "config4": dict(
        ps=dict(COMMON, rate_diffusion=0.005, rate_active=10, init="poisson", N=500, local_kernel_sigma=0.02), # double_sweep.py:674-694
        run=dict(T=10, obs_dt=0.1, record_fft=False, record_var=False),                                        # :666-671
        profile=dict(frac_plus=0.75, decay_plus=0.2, decay_minus=0.2), scaling="strong",
        text="BASELINE config 4: (density, beta) double sweep, 16 densities (N = 50..950) x 16 beta x 128 replicas = 32 768 "
             "replicas in total at every N, L=1000, K=1, sigma=0.02 (r=80), D=0.005, lambda=10, run(T=10, obs_dt=0.1)"),
}
K2_TEXT = ("BASELINE config 5: one lattice of 2^26 sites in total (slabs over the GPUs, in-kernel NVLink ghost exchange), K=1, "
           "density 0.5, local Gaussian field sigma = 5 sites (r=20), D=0.02, lambda=5, beta=2, dt=0.005 (2 passes per dt)")


def init_kwargs(wl="config2"):
    from aps_b200.launcher import make_exp_gradient
    p = WORKLOADS[wl]["profile"]
    return dict(rho0_plus=make_exp_gradient(L=1000, N=p["N"], frac_plus=p["frac_plus"], decay_length=p["decay_plus"], anchor_positions=None)[0],
                rho0_minus=make_exp_gradient(L=1000, N=p["N"], frac_plus=p["frac_plus"], decay_length=p["decay_minus"], anchor_positions=None)[1])


PS_KWARGS, RUN_KWARGS = WORKLOADS["config2"]["ps"], WORKLOADS["config2"]["run"]      # used by tools/*.py


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (every 250 ms: polling perturbs the running
    kernels measurably — 100 ms polling cost 1.6 % of the step on a B200)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_ev, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "250"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._stop_ev.is_set():
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_ev.set()
        if self.proc:
            self.proc.terminate()
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower() == "active"})
        return dict(sm_mhz=int(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def hbm_peak():
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        return json.load(open(peaks))["hbm_gbs"], "MEASURED_PEAKS.json (measured copy bandwidth)"
    return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------- CPU arms ----------------------------------------------
def sample_points(wl, n):
    """`n` (beta, N_part) points spread over the workload's sweep grid."""
    betas = np.linspace(0, 3, 64 if wl != "config4" else 16)
    b = betas[np.linspace(0, len(betas) - 1, n).round().astype(int)] if n < len(betas) else np.resize(betas, n)
    if wl == "config4":
        dens = np.linspace(50, 950, 16).astype(int)
        N = np.resize(dens[::3], n)
    else:
        N = np.full(n, WORKLOADS[wl]["ps"]["N"])
    return [(float(x), int(y)) for x, y in zip(b, N)]


_REF = {}


def _ref_worker(task):
    """One replica through the UNMODIFIED reference (baseline/_ref): returns (events, seconds).  Runs in a worker process."""
    wl, beta, n_part, seed, T = task
    if "PS" not in _REF:
        import plot_stubs
        plot_stubs.install()                      # the reference imports matplotlib / vispy at module top (plot-only)
        sys.path.insert(0, REF_DIR)
        from PARTICLE_solver_CLASS import ParticleSystem
        src = open(os.path.join(REF_DIR, "PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py")).read()
        ns = {"np": np}                           # the driver's own make_exp_gradient (sweep_beta.py:16-53), cut out by AST:
        for node in ast.parse(src).body:          # the script itself runs its sweep at import time
            if isinstance(node, ast.FunctionDef) and node.name == "make_exp_gradient":
                exec(compile(ast.Module(body=[node], type_ignores=[]), "sweep_beta.py", "exec"), ns)
        _REF.update(PS=ParticleSystem, grad=ns["make_exp_gradient"])
    w = WORKLOADS[wl]

    class Counting:                               # forwards every call; one exponential per event (CLASS.py:358)
        def __init__(self, g): self.g, self.n = g, 0
        def exponential(self, *a, **k): self.n += 1; return self.g.exponential(*a, **k)
        def choice(self, *a, **k): return self.g.choice(*a, **k)
        def random(self, *a, **k): return self.g.random(*a, **k)
        def poisson(self, *a, **k): return self.g.poisson(*a, **k)

    kw = dict(w["ps"], beta=beta, N=n_part)
    if w["profile"] is not None:
        p = w["profile"]
        kw["rho0_plus"] = _REF["grad"](L=1000, N=n_part, frac_plus=p["frac_plus"], decay_length=p["decay_plus"], anchor_positions=None)[0]
        kw["rho0_minus"] = _REF["grad"](L=1000, N=n_part, frac_plus=p["frac_plus"], decay_length=p["decay_minus"], anchor_positions=None)[1]
    rng = Counting(np.random.default_rng(seed))
    ps = _REF["PS"](rng=rng, **kw)
    t0 = time.perf_counter()
    ps.run(**dict(w["run"], T=T))
    return rng.n, time.perf_counter() - t0


class ReferencePool:
    """nproc worker processes, each running the unmodified numpy reference (spawned: no CUDA state is inherited)."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores)
        self.pool.map(_ref_worker, [("config2", 1.0, 500, i, 0.05) for i in range(2 * cores)], chunksize=1)   # imports, untimed

    def step(self, wl, T, seed0):
        pts = sample_points(wl, self.cores)
        t0 = time.perf_counter()
        res = self.pool.map(_ref_worker, [(wl, b, n, seed0 * 100_003 + i, T) for i, (b, n) in enumerate(pts)], chunksize=1)
        return sum(r[0] for r in res), time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "PARTICLE_solver_CLASS.py"))


def oracle_sample(wl, n_replicas, T, threads, seed0=0):
    """Port arm: `n_replicas` replicas of the workload through the C restatement of the reference algorithm (oracle/),
    `threads` at a time.  Returns (events, seconds)."""
    from aps_b200 import launcher as la
    from aps_b200.batch import make_params
    from common import HostRun
    from oracle import oracle
    from oracle_ensemble import OracleEnsemble
    w = WORKLOADS[wl]
    pts = sample_points(wl, n_replicas)
    if wl == "config4":
        p = w["profile"]
        spec = la.build_double_sweep_spec(sorted({n for _, n in pts}), [0.0], 1, w["ps"], dict(w["run"], T=T), frac_plus=p["frac_plus"],
                                          decay_plus=p["decay_plus"], decay_minus=p["decay_minus"], base_seed=seed0)
        order = {n: i for i, n in enumerate(sorted({n for _, n in pts}))}
        spec.betas = np.array([b for b, _ in pts]); spec.profile_of = np.array([order[n] for _, n in pts], np.int32)
        spec.point_of = np.arange(len(pts)); spec.seeds = np.arange(len(pts), dtype=np.uint64) + np.uint64(1000 * seed0)
    else:
        spec = la.build_beta_sweep_spec([b for b, _ in pts], 1, w["ps"], init_kwargs(wl) if w["profile"] else {}, dict(w["run"], T=T),
                                        base_seed=seed0)
    ens = OracleEnsemble(spec, 0, n_replicas)
    seeds, pos0, sg0, n = ens.init_states()
    mp_ = ens.mp
    hr = HostRun(mp_["L"], ens.n_max, len(ens.times_obs), n, pos0, sg0, spec.betas, ens.times_obs, mp_["weights"], seeds=seeds,
                 record=3, alloc_m_local=False)
    P = make_params(mp_["L"], mp_["K"], mp_["radius"], mp_["D"], mp_["lam"], float(T))
    t0 = time.perf_counter()
    assert oracle.load().aps_oracle_run(P, hr.batch, 1, threads) == 0
    return int(hr.n_events.sum()), time.perf_counter() - t0


def cpu_arm(wl, budget_s, steps=None, warmup=0):
    """Times the CPU implementation on a bounded sample.  steps=None: as many steps as fit `budget_s` (cpu_baseline leg);
    otherwise exactly `steps` timed steps after `warmup` (reference arm).  Returns the cpu_baseline dict + ms per step."""
    cores = os.cpu_count() or 1
    out = {}
    if have_reference():
        T = 2.0
        pool = ReferencePool(cores)
        try:
            ev = dt = 0.0
            k = 0
            n_steps = 0
            while True:
                e, d = pool.step(wl, T, seed0=k)
                k += 1
                if k > warmup:
                    ev, dt, n_steps = ev + e, dt + d, n_steps + 1
                if (steps is not None and n_steps >= steps) or (steps is None and dt > budget_s):
                    break
        finally:
            pool.close()
        out = dict(value=ev / dt, unit=UNIT, cores=cores, kind="reference", ms_per_step=1e3 * dt / n_steps, steps=n_steps,
                   sample=f"{cores} replicas per step (one per core, betas spread over the sweep), run(T={T}) instead of the full "
                          f"length, the UNMODIFIED numpy ParticleSystem.run from baseline/_ref in {cores} processes")
    n_rep = min(256, 4 * cores)
    ev = dt = 0.0
    n_steps = 0
    port_budget = budget_s if not out else min(6.0, budget_s)
    while True:
        e, d = oracle_sample(wl, n_rep, 5.0, cores, seed0=n_steps)
        ev, dt, n_steps = ev + e, dt + d, n_steps + 1
        if (out and dt > port_budget) or (not out and ((steps is not None and n_steps >= steps + warmup) or (steps is None and dt > budget_s))):
            break
    port = dict(value=ev / dt, unit=UNIT, cores=cores, kind="port", ms_per_step=1e3 * dt / n_steps, steps=n_steps,
                sample=f"{n_rep} replicas per step, run(T=5), {cores} pthreads, C restatement of the reference algorithm "
                       "(full field + all rates per event, oracle/aps_oracle.c)")
    if out:
        out["port"] = port
        return out
    return port


def k2_cpu_arm(budget_s):
    """K2 has no reference counterpart; CPU arm = the oracle's sequential restatement of the same update rule."""
    from aps_b200.sublattice import SublatticeLattice
    from oracle_k2 import OracleK2Backend
    L = 1 << 20
    lat = SublatticeLattice(L, D=0.02, lam=5.0, beta=2.0, dt=0.005, sigma_sites=5.0, seed=0, backend=OracleK2Backend(), single_rank=True)
    lat.init_random(0.5, 0.5)
    lat.run_passes(2)
    passes, dt = 0, 0.0
    while dt < budget_s:
        t0 = time.perf_counter(); lat.run_passes(4); dt += time.perf_counter() - t0; passes += 4
    return dict(value=lat.n_particles * passes / dt, unit=UNIT, cores=1, kind="port", ms_per_step=1e3 * dt / passes,
                sample=f"lattice of 2^20 sites, {passes} passes, single thread, oracle restatement of the sublattice update rule "
                       "(the reference itself cannot run lattices of this kind: O(L) numpy work per event)")


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload
    t0 = time.perf_counter()
    cb = k2_cpu_arm(8.0 * max(1, args.steps) / 5) if wl == "k2" else cpu_arm(wl, 10.0, steps=args.steps, warmup=args.warmup)
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=cb.get("ms_per_step"), higher_is_better=True,
                scaling="strong" if wl in ("config4", "k2") else "weak", vs_baseline=None, dtype="f64" if wl != "k2" else "u8",
                data="synthetic", impl="reference", config=dict(workload=K2_TEXT if wl == "k2" else WORKLOADS[wl]["text"]),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=round(time.perf_counter() - t0, 1))
    print(json.dumps(line))


# ---------------------------------------------------------------- K2 measurements -----------------------------------------
def k2_case(torch, L, sigma, dt, passes, persistent=None, single_rank=True):
    from aps_b200.sublattice import SublatticeLattice
    lat = SublatticeLattice(L, D=0.02, lam=5.0, beta=2.0, dt=dt, sigma_sites=sigma, seed=0, single_rank=single_rank, persistent=persistent)
    lat.init_random(0.5, 0.5)
    lat.run_passes(6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lat.run_passes(passes); e1.record(); torch.cuda.synchronize()
    lat.check()
    ms = e0.elapsed_time(e1) / passes
    res = dict(ms_per_pass=ms, particle_attempts_per_s=lat.n_particles / (ms * 1e-3), trials_per_half_pass=lat.rates.mu,
               simulated_time_per_s=0.5 * dt / (ms * 1e-3))
    lat.close()
    del lat
    torch.cuda.empty_cache()
    return res


def measure_k2(torch):
    """K2 roofline on ONE GPU at the operating point BASELINE config 5 names (L = 2^26, local field sigma = 5 sites at
    dt = 0.005; global field at dt = 0.02, the largest dt with lambda*dt <= 0.1) and, for reference, at L = 2^30 and at smaller
    dt.  achieved = algorithmic 2 B per site-visit (read 1 B + write 1 B) x sites / pass time, against the measured copy peak.
    The HEADLINE is the config-5 case itself (local field, dt = 0.005, L = 2^26), not the best case."""
    peak, src = hbm_peak()
    out = dict(bound="hbm", kernel="aps::k2_pass_kernel", unit="GB/s", peak=peak, peak_source=src,
               algorithmic_bytes_per_site_visit=2, cases=[])
    traffic_file = os.path.join(ROOT, "profiles", "k2_ncu_traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    for logL, passes in [(26, 200), (30, 24)]:
        for name, sigma, dt in [("local Gaussian field sigma=5 sites, dt=0.005", 5.0, 0.005), ("global field, dt=0.02", None, 0.02),
                                ("local Gaussian field sigma=5 sites, dt=0.0025", 5.0, 0.0025), ("global field, dt=0.005", None, 0.005),
                                ("global field, dt=0.0025", None, 0.0025)]:
            r = k2_case(torch, 1 << logL, sigma, dt, passes)
            gbs = 2.0 * (1 << logL) / (r["ms_per_pass"] * 1e-3) / 1e9
            out["cases"].append(dict(case=f"L=2^{logL}, {name}", achieved=gbs, frac=gbs / peak, traffic=traffic.get(name), **r))
    head = out["cases"][0]
    out.update(achieved=head["achieved"], frac=head["frac"], traffic=head["traffic"], headline_case=head["case"],
               simulated_time_per_s=head["simulated_time_per_s"])
    return out


# ---------------------------------------------------------------- GPU arms ------------------------------------------------
def build_spec(la, wl, world, reps_scale=1):
    w = WORKLOADS[wl]
    if wl == "config4":
        p = w["profile"]
        return la.build_double_sweep_spec(np.linspace(50, 950, 16).astype(int), np.linspace(0, 3, 16), 128, w["ps"], w["run"],
                                          frac_plus=p["frac_plus"], decay_plus=p["decay_plus"], decay_minus=p["decay_minus"], base_seed=1)
    reps = REPS_PER_BETA * (world if w["scaling"] == "weak" else 1) * reps_scale
    spec = la.build_beta_sweep_spec(np.linspace(0, 3, N_BETA), reps, w["ps"], init_kwargs(wl) if w["profile"] else {}, w["run"], base_seed=1)
    if wl == "config3":
        from aps_b200.capi import APS_REC_COUNTS, APS_REC_MLOCAL, APS_REC_POS
        spec.record = APS_REC_COUNTS | APS_REC_POS | APS_REC_MLOCAL
    return spec


def ensemble_main(args, torch, la, capi, rank, world):
    wl = args.workload
    w = WORKLOADS[wl]
    lib = capi.load()
    dev = torch.device("cuda", torch.cuda.current_device())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def make_ens(spec):
        lo, hi = la.shard_bounds(len(spec.betas), rank, world)
        return la.DeviceEnsemble(la.permute_spec(spec, la.schedule_order(spec, world)), lo, hi)   # the launcher's own schedule

    if wl == "config3":
        from aps_b200.structure import fft_amplitudes, structure_observables

        def step(ens, ev=None):                # init -> K1 -> density rows, cuFFT amplitudes, per-run structure observables
            ens.init_particles()
            if ev is not None:
                ev[0].record()
            ens.rb.run_philox()
            if ev is not None:
                ev[1].record()
            amp, total, var = fft_amplitudes(ens.rb)
            ens.struct = structure_observables(ens.rb, 0.5, None, amp=amp, var=var)
    else:
        def step(ens):                         # init -> K1 -> reducers, histogram, per-point profile sums (hand-written kernels only)
            ens.step(want_profiles=(wl == "config2"))

    # ---------------- device-resident arm (value) ----------------
    ens = make_ens(build_spec(la, wl, world))
    W = max(3, args.warmup)
    for _ in range(W):
        step(ens)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    if not os.environ.get("APS_BENCH_NO_SAMPLER"):
        sampler.start()
    time.sleep(0.3)
    n0 = lib.aps_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    e0.record()
    for s in range(args.steps):                # every step re-runs the same seeded ensemble from its initial conditions
        if wl == "config3":
            step(ens, k_ev[s])
        else:
            ens.init_particles()
            k_ev[s][0].record()
            ens.rb.run_philox()
            k_ev[s][1].record()
            ens.red = ens.rb.reduce()
            ens.hist, ens.mbar = ens.rb.m_histogram(max(1, ens.n_points), ens.point_of)
            if wl == "config2":
                ens.prof = ens.rb.profile_sums_by_point(ens.n_points, ens.point_start, ens.point_reps)
    e1.record()
    barrier()
    launches = lib.aps_launch_count() - n0
    k1_ms = sum(a.elapsed_time(b) for a, b in k_ev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    evs = (ens.rb.n_events.sum().double() * args.steps).reshape(1)       # read after the timed region (identical steps)
    events_per_launch = float(evs.item()) / args.steps
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(evs, op=torch.distributed.ReduceOp.SUM)
    total_ms, total_events = float(ms.item()), float(evs.item())
    value = total_events / (total_ms * 1e-3)
    mean_n = float(ens.n.double().mean().item())
    guard = int(ens.rb.n_guard.sum().item())
    bad = int((ens.rb.status != 0).sum().item())
    n_rep_gpu = ens.R
    del ens
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm through the public API (host parameters -> host results) ----------------
    if wl == "config2":
        betas, reps, ik = np.linspace(0, 3, N_BETA), REPS_PER_BETA * world, init_kwargs(wl)
        call = lambda s: la.sweep_over_betas(betas, reps, w["ps"], ik, w["run"], base_seed=100 + s)
        count = lambda out: (int(out["n_events"].sum()), out["info"]["h2d_bytes"], out["info"]["d2h_bytes"])
        api = "launcher.sweep_over_betas (host parameters -> host reducers, per-beta profiles, magnetisation histogram)"
    elif wl == "config3":
        betas, reps = np.linspace(0, 3, N_BETA), REPS_PER_BETA
        betas_all = np.linspace(0, 3, N_BETA * world) if world > 1 else betas          # weak: 64 beta values per GPU
        call = lambda s: la.sweep_betas_for_structures(list(betas_all), reps, w["ps"], {}, w["run"], base_seed=100 + s, keep_raw=False)
        count = lambda out: (int(sum(int(v["n_events"].sum()) for v in out.values())), 8 * 64 * 3, sum(8 * (v["fft_mean_mean"].size * 2 + 16) for v in out.values()))
        api = "launcher.sweep_betas_for_structures (host parameters -> host structure observables per beta)"
    else:
        p = w["profile"]
        dens, betas = np.linspace(50, 950, 16).astype(int), np.linspace(0, 3, 16)
        call = lambda s: la.double_sweep(dens, betas, 128, w["ps"], w["run"], frac_plus=p["frac_plus"], decay_plus=p["decay_plus"],
                                         decay_minus=p["decay_minus"], base_seed=100 + s)
        count = lambda out: (int(out["info"]["n_events_total"]), out["info"]["h2d_bytes"], out["info"]["d2h_bytes"])
        api = "launcher.double_sweep (host parameters -> host reducers per (density, beta) point)"
    for _ in range(2):
        call(0)
    barrier()
    t0 = time.perf_counter()
    e2e_events, h2d, d2h = 0, 0, 0
    for s in range(args.steps):
        ev, h2d, d2h = count(call(s))
        e2e_events += ev
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_value = e2e_events / float(e2e_s.item())      # event counts are already gathered over all ranks
    clocks = sampler.stop()

    # ---------------- config 2 at N > 1: the 4096-replica sweep itself sharded over the GPUs (strong scaling) ----------------
    strong = None
    if wl == "config2" and world > 1:
        spec = la.build_beta_sweep_spec(np.linspace(0, 3, N_BETA), REPS_PER_BETA, w["ps"], init_kwargs(wl), w["run"], base_seed=1)
        e2 = make_ens(spec)
        for _ in range(2):
            step(e2)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            step(e2)
        a1.record()
        barrier()
        t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        ev = (e2.rb.n_events.sum().double() * 3).reshape(1)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(ev, op=torch.distributed.ReduceOp.SUM)
        strong = dict(total_replicas=N_BETA * REPS_PER_BETA, replicas_per_gpu=e2.R, value=float(ev.item()) / (float(t.item()) * 1e-3),
                      ms_per_step=float(t.item()) / 3,
                      note="BASELINE config 2 as named (4096 replicas in total): a launch lasts as long as its slowest replica, so fewer "
                           "replicas per GPU do not shorten it — the exact chain does not strong-scale below one wave of CTAs")
        del e2

    k2 = None
    if rank == 0 and wl == "config2" and not args.no_k2:
        try:
            k2 = measure_k2(torch)
        except Exception as exc:            # never let the secondary measurement break the contract line
            k2 = dict(error=str(exc)[:300])

    if rank == 0:
        roofline = None
        if k1_ms is not None:
            # K1 roofline: shared-memory bandwidth (the lattice never leaves the SM)
            r, L = (80 if wl == "config4" else 20), 1000
            bytes_per_event = 8 * mean_n * 2 + (mean_n / L) * (2 * r + 2) * (r + 1) * (2 * 2 + 16)
            k1_avg_ms = k1_ms / args.steps
            sm_mhz = (clocks.get("sm_max_mhz") or 1965)
            peak = 148 * 128 * sm_mhz * 1e6 / 1e9
            peak_source = ("computed 148 SM x 128 B/clk x clocks.max.sm (shared-memory bandwidth is not in MEASURED_PEAKS.json); "
                           "HBM traffic of K1 is only the observation rows")
            sp_file = os.path.join(ROOT, "profiles", "smem_peak.json")
            if os.path.exists(sp_file):        # measured on a B200 of this pool with tools/smem_peak.py (LDS.128 microbenchmark): 99.8 % of the formula
                sp = json.load(open(sp_file))
                peak = float(sp["smem_read_gbs"])
                peak_source = ("MEASURED shared-memory read bandwidth, tools/smem_peak.py -> profiles/smem_peak.json (%.1f B/clk/SM at the max clock; "
                               "MEASURED_PEAKS.json has no shared-memory figure; formula 148 x 128 B/clk x 1965 MHz = 37 225 GB/s); "
                               "HBM traffic of K1 is only the observation rows" % sp["bytes_per_clk_per_sm_at_max_clock"])
            achieved = bytes_per_event * events_per_launch / (k1_avg_ms * 1e-3) / 1e9
            ncu_file = os.path.join(ROOT, "profiles", "k1_ncu_summary.json")
            ncu_k1 = json.load(open(ncu_file)) if os.path.exists(ncu_file) else {}
            smem_traffic = ncu_k1.get("smem_bytes_per_event_at_128B_per_wavefront") if wl == "config2" else None
            kname = {"config2": "aps::k1_lean_kernel<true,21,1056,512,false>", "config3": "aps::k1_lean_kernel<true,21,1056,1024,false>",
                     "config4": "aps::k1_lean_kernel<true,81,1184,{512,1024},false>"}[wl]
            roofline = dict(bound="smem", kernel=kname,
                            achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                            traffic=(smem_traffic * events_per_launch if smem_traffic else None),
                            traffic_note="shared-memory wavefronts x 128 B per launch from the ncu capture (DRAM traffic of K1 is ~0.1 GB per launch)",
                            peak_source=peak_source,
                            algorithmic_bytes_per_event=bytes_per_event, events_per_launch=events_per_launch, kernel_ms=k1_avg_ms,
                            kernel_share_of_step=k1_avg_ms * args.steps / total_ms, ncu=ncu_k1 if wl == "config2" else None)
        cpu_baseline = None if args.no_cpu_baseline else cpu_arm(wl, 12.0)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=W, ms_per_step=total_ms / args.steps,
                    higher_is_better=True, scaling=w["scaling"], vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=w["text"], replicas_per_gpu=n_rep_gpu, mode="native Philox4x32-10, device-side init",
                                cache="per-step working set (observation rows, GBs) exceeds the 126 MB L2; no flush needed",
                                mean_particles=mean_n, events_per_step_per_gpu=events_per_launch, guard_fallbacks=guard,
                                replicas_not_done=bad, strong_scaling_of_the_4096_replica_sweep=strong),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), api=api),
                    gpu_launches=int(launches), clocks=clocks, roofline=roofline, roofline_k2=k2, cpu_baseline=cpu_baseline)
        print(json.dumps(line))


def k2_main(args, torch, la, capi, rank, world):
    """BASELINE config 5: one lattice of 2^26 sites cut into slabs over the GPUs (strong scaling), persistent kernel with the
    in-kernel NVLink exchange.  A step = 2 passes (one time unit dt).  e2e: host lattice -> H2D -> passes -> coarse profile -> host."""
    from aps_b200.sublattice import SublatticeLattice
    lib = capi.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    L, sigma, dt, passes_per_step = 1 << 26, 5.0, 0.005, 2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(lat, steps, warm):
        lat.run_passes(passes_per_step * warm)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lat.run_passes(passes_per_step * steps); e1.record()
        barrier()
        lat.check()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    steps = max(args.steps, 50)                # a step lasts ~0.1 ms: time at least 50 of them
    W = max(3, args.warmup)
    lat = SublatticeLattice(L, D=0.02, lam=5.0, beta=2.0, dt=dt, sigma_sites=sigma, seed=0)
    lat.init_random(0.5, 0.5)
    sampler = ClockSampler(dev.index or 0)
    if not os.environ.get("APS_BENCH_NO_SAMPLER"):
        sampler.start()
    n0 = lib.aps_launch_count()
    total_ms = timed(lat, steps, W)
    launches = lib.aps_launch_count() - n0
    value = lat.n_particles * passes_per_step * steps / (total_ms * 1e-3)
    peak, src = hbm_peak()
    gbs = 2.0 * L * passes_per_step * steps / (total_ms * 1e-3) / 1e9
    extra = []
    for name, LL, sg, dtt in [("L=2^30 (strong), local field sigma=5, dt=0.005", 1 << 30, 5.0, 0.005),
                              ("L=2^26 (strong), global field, dt=0.02", 1 << 26, None, 0.02),
                              ("L=2^30 (strong), global field, dt=0.02", 1 << 30, None, 0.02)]:
        l2 = SublatticeLattice(LL, D=0.02, lam=5.0, beta=2.0, dt=dtt, sigma_sites=sg, seed=0)
        l2.init_random(0.5, 0.5)
        st = 50 if LL <= (1 << 26) else 10
        ms = timed(l2, st, 2)
        extra.append(dict(case=name, ms_per_pass=ms / (st * passes_per_step), particle_attempts_per_s=l2.n_particles * passes_per_step * st / (ms * 1e-3),
                          GBs_all_gpus=2.0 * LL * passes_per_step * st / (ms * 1e-3) / 1e9))
        l2.close()
        del l2
        torch.cuda.empty_cache()
    # ---- end to end: host lattice bytes -> device slabs -> passes -> coarse profile on the host ----
    rng = np.random.default_rng(0)
    host = torch.from_numpy(np.where(rng.random(L) < 0.5, np.where(rng.random(L) < 0.5, 1, 2), 0).astype(np.uint8)).pin_memory()
    e2e_steps = 100                          # upload once, run 100 time units, read the profile: the shape of a real use
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        lat.set_state_from_host(host)          # H2D of this rank's slab (pinned)
        lat.run_passes(passes_per_step * e2e_steps)
        prof = lat.profile(1000)               # D2H of the coarse profile (all-reduced counts)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_value = 3 * lat.n_particles * passes_per_step * e2e_steps / float(e2e_s.item())
    clocks = sampler.stop()
    slab = lat.L
    lat.close()
    if rank == 0:
        cpu_baseline = None if args.no_cpu_baseline else k2_cpu_arm(8.0)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=W, ms_per_step=total_ms / steps,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="u8", data="synthetic",
                    config=dict(workload=K2_TEXT, sites_per_gpu_incl_ghosts=slab, passes_per_step=passes_per_step,
                                launch="persistent cooperative kernel, grid barrier per pass, ghost refresh through CUDA-IPC peer memory" if world > 1
                                       else "one launch per pass (a grid barrier per pass measures 5-10 % slower on one GPU)",
                                cache="2 x 64 MiB ping-pong buffers per lattice: comparable to the 126 MB L2 (config 5 is that size); L = 2^30 cases below are 8x the L2",
                                other_cases=extra),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(slab / e2e_steps), d2h_bytes_per_step=int(2 * 1000 * 8 / e2e_steps),
                             api=f"SublatticeLattice.set_state_from_host / run_passes / profile: host lattice in, coarse profile out, per {e2e_steps} steps"),
                    gpu_launches=int(launches), clocks=clocks,
                    roofline=dict(bound="hbm", kernel="aps::k2_pass_kernel<true,*>", achieved=gbs, peak=peak * world, unit="GB/s", frac=gbs / (peak * world),
                                  traffic=None, peak_source=src + f" x {world} GPUs", algorithmic_bytes_per_site_visit=2),
                    cpu_baseline=cpu_baseline)
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "k2"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-k2", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from aps_b200 import capi, launcher as la

    rank, world = la.init_distributed_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if args.workload == "k2":
        k2_main(args, torch, la, capi, rank, world)
    else:
        ensemble_main(args, torch, la, capi, rank, world)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
