#!/usr/bin/env python
"""Benchmark contract (see the task statement).

Metric (BASELINE.json): particle update attempts / second over a beta-sweep ensemble.
Workload at N=1: BASELINE config 2 — PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta parameters
(sweep_beta.py:829-878), 64 beta points x 64 replicas = 4096 independent ParticleSystem.run(T=20,
obs_dt=0.1) calls.  One exact-Gillespie event == one particle update attempt (every event is an
accepted update, CLASS.py:351-367).  Weak scaling: every rank runs 4096 replicas (64 beta x 64N).

A step = device init of the 4096 replicas (K3) -> K1 time stepping -> K4 per-run reducers and
per-beta profile sums.  `value` times that with inputs resident in HBM (CUDA events); `e2e` times
the public call `launcher.sweep_over_betas(...)` from host parameters to host results (wall clock
around a call that ends in a device->host copy), H2D and D2H included.

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port":
the reference is Python and cannot travel to the GPU box) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
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

PS_KWARGS = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, flip_rate_fn=None, init="poisson", N=500,
                 scale_rates=False, local_kernel_sigma=0.005, minus_anchor=True, periodic=False,
                 immobilize_when_anchored=True, anchor_radius=0.003, anchor_positions=None, site_capacity=1,
                 crowding_suppresses_rates=False, k_on=0, k_off=0, k_exit=0)          # sweep_beta.py:837-857
RUN_KWARGS = dict(T=20, obs_dt=0.1, record_fft=True, record_var=True)                 # sweep_beta.py:829-834


def init_kwargs():
    from aps_b200.launcher import make_exp_gradient
    return dict(rho0_plus=make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)[0],
                rho0_minus=make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.2, anchor_positions=None)[1])


CONFIG = dict(workload="BASELINE config 2: sweep_beta ensemble, 64 beta x 64 replicas, L=1000, Poisson init N=500, "
                       "K=1, sigma=0.005 (r=20), D=0.02, lambda=5, run(T=20, obs_dt=0.1)",
              replicas_per_gpu=N_BETA * REPS_PER_BETA, mode="native Philox4x32-10, device-side init",
              cache="per-step working set (observation rows, 3.4 GB) exceeds the 126 MB L2; no flush needed")


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


def oracle_sample(n_replicas, T, threads, seed0=0):
    """CPU arm: `n_replicas` replicas of the workload (betas spread over the sweep) through the oracle,
    `threads` replicas at a time.  Returns (events, seconds)."""
    from aps_b200.launcher import build_beta_sweep_spec
    from oracle_ensemble import OracleEnsemble
    betas = np.linspace(0, 3, N_BETA)[np.linspace(0, N_BETA - 1, n_replicas).astype(int)] if n_replicas < N_BETA \
        else np.resize(np.linspace(0, 3, N_BETA), n_replicas)
    spec = build_beta_sweep_spec(betas, 1, PS_KWARGS, init_kwargs(), dict(RUN_KWARGS, T=T), base_seed=seed0)
    ens = OracleEnsemble(spec, 0, n_replicas)
    seeds, pos0, sg0, n = ens.init_states()
    from aps_b200.batch import make_params
    from common import HostRun
    from oracle import oracle
    mp = ens.mp
    hr = HostRun(mp["L"], ens.n_max, len(ens.times_obs), n, pos0, sg0, spec.betas, ens.times_obs, mp["weights"],
                 seeds=seeds, record=3)
    P = make_params(mp["L"], mp["K"], mp["radius"], mp["D"], mp["lam"], float(T))
    lib = oracle.load()
    t0 = time.perf_counter()
    assert lib.aps_oracle_run(P, hr.batch, 1, threads) == 0
    dt = time.perf_counter() - t0
    return int(hr.n_events.sum()), dt


def measure_k2(torch, logL=30, passes=30):
    """K2 on one lattice of 2^30 sites (1 GiB per buffer, far larger than the 126 MB L2): achieved HBM GB/s =
    algorithmic 2 B per site-visit (read 1 B + write 1 B) x sites / pass time, against the measured copy peak."""
    from aps_b200.sublattice import SublatticeLattice
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = (json.load(open(peaks))["hbm_gbs"], "MEASURED_PEAKS.json (measured)") if os.path.exists(peaks) else (6650.0, "fallback")
    out = dict(bound="hbm", kernel="aps::k2_pass_kernel", unit="GB/s", peak=peak, peak_source=src, L=1 << logL,
               algorithmic_bytes_per_site_visit=2, cases=[])
    traffic_file = os.path.join(ROOT, "profiles", "k2_ncu_traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    for name, sigma, dt in [("global field, dt=0.0025", None, 0.0025), ("global field, dt=0.005", None, 0.005),
                            ("global field, dt=0.02", None, 0.02), ("local Gaussian field sigma=5 sites, dt=0.005", 5.0, 0.005),
                            ("local Gaussian field sigma=5 sites, dt=0.0025", 5.0, 0.0025)]:
        lat = SublatticeLattice(1 << logL, D=0.02, lam=5.0, beta=2.0, dt=dt, sigma_sites=sigma, seed=0,
                                single_rank=True)     # the K2 roofline is a one-GPU measurement on rank 0 at every N
        lat.init_random(0.5, 0.5)
        lat.run_passes(6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lat.run_passes(passes); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / passes
        gbs = 2.0 * (1 << logL) / (ms * 1e-3) / 1e9
        out["cases"].append(dict(case=name, ms_per_pass=ms, achieved=gbs, frac=gbs / peak,
                                 particle_attempts_per_s=lat.n_particles / (ms * 1e-3), trials_per_segment_pass=lat.rates.mu,
                                 traffic=traffic.get(name)))
        del lat
        torch.cuda.empty_cache()
    best = max(out["cases"], key=lambda c: c["frac"])
    out.update(achieved=best["achieved"], frac=best["frac"], traffic=best["traffic"], headline_case=best["case"])
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_rep = min(256, 4 * cores)
    T = 5.0
    times, events = [], 0
    for s in range(args.warmup + args.steps):
        ev, dt = oracle_sample(n_rep, T, cores, seed0=s)
        if s >= args.warmup:
            times.append(dt); events += ev
    value = events / sum(times)
    sample = f"{n_rep} of the 4096 replicas per step (betas spread over the sweep), run(T={T}) instead of T=20, " \
             f"{cores} pthreads, C restatement of the reference algorithm (full field + all rates per event)"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference", config=CONFIG,
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="the reference's own numpy path measured 3.4-4.1e3 events/s/core in the build container "
                     "(BASELINE.md section 2); it cannot run on the GPU box")
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--T", type=float, default=float(RUN_KWARGS["T"]), help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from aps_b200 import capi, launcher as la

    rank, world = la.init_distributed_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    lib = capi.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    run_kwargs = dict(RUN_KWARGS, T=args.T)
    ik = init_kwargs()
    reps = REPS_PER_BETA * world
    betas = np.linspace(0, 3, N_BETA)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (value) ----------------
    spec = la.build_beta_sweep_spec(betas, reps, PS_KWARGS, ik, run_kwargs, base_seed=1)
    lo, hi = la.shard_bounds(len(spec.betas), rank, world)
    # same schedule as launcher.run_ensemble: every rank gets 64 replicas of every beta, longest-running first
    spec = la.permute_spec(spec, la.schedule_order(spec, world))
    ens = la.DeviceEnsemble(spec, lo, hi)
    for _ in range(max(3, args.warmup)):
        ens.step()
    barrier()
    sampler = ClockSampler(dev.index or 0)
    if not os.environ.get("APS_BENCH_NO_SAMPLER"):
        sampler.start()
    time.sleep(0.3)
    n0 = lib.aps_launch_count()
    ev_total = torch.zeros((), dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    e0.record()
    for s in range(args.steps):
        ens.init_particles()
        k_ev[s][0].record()
        ens.rb.run_philox()
        k_ev[s][1].record()
        ens.red = ens.rb.reduce()
        per_rep = ens.rb.profile_sums(1)
        ens.prof = torch.zeros((ens.n_points, 4, 1000), dtype=torch.float64, device=dev).index_add_(0, ens.point_local, per_rep)
        ev_total += ens.rb.n_events.sum()
    e1.record()
    barrier()
    k1_ms = sum(a.elapsed_time(b) for a, b in k_ev)      # K1 launches of the timed steps (events read after the final sync)
    launches = lib.aps_launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if os.environ.get("APS_BENCH_DEBUG"):
        print(f"[rank {rank}] device-arm ms/step {float(ms) / args.steps:.2f}  K1 {sum(a.elapsed_time(b) for a, b in k_ev) / args.steps:.2f}", file=sys.stderr, flush=True)
    evs = ev_total.double().reshape(1)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(evs, op=torch.distributed.ReduceOp.SUM)
    total_ms, total_events = float(ms.item()), float(evs.item())
    value = total_events / (total_ms * 1e-3)
    events_per_launch = float(ev_total.item()) / args.steps
    mean_n = float(ens.n.double().mean().item())
    guard = int(ens.rb.n_guard.sum().item())
    bad = int((ens.rb.status != 0).sum().item())

    # ---------------- end-to-end arm through the public API (host -> host) ----------------
    for _ in range(2):
        la.sweep_over_betas(betas, reps, PS_KWARGS, ik, run_kwargs, base_seed=2)
    barrier()
    t0 = time.perf_counter()
    e2e_events, info = 0, None
    for s in range(args.steps):
        out = la.sweep_over_betas(betas, reps, PS_KWARGS, ik, run_kwargs, base_seed=100 + s)
        e2e_events += int(out["n_events"].sum())
        info = out["info"]
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    clocks = sampler.stop()
    e2e_value = e2e_events / float(e2e_s.item())      # n_events is already gathered over all ranks

    # ---------------- K2 (sublattice kernel, HBM-bound) measured beside the main workload ----------------
    k2 = None
    if rank == 0:
        try:
            k2 = measure_k2(torch)
        except Exception as exc:            # never let the secondary measurement break the contract line
            k2 = dict(error=str(exc)[:200])

    if rank == 0:
        # K1 roofline: shared-memory bandwidth (the lattice never leaves the SM)
        r, L = 20, 1000
        bytes_per_event = 8 * mean_n * 2 + (mean_n / L) * (2 * r + 2) * (r + 1) * (2 * 2 + 16)
        k1_avg_ms = k1_ms / args.steps
        sm_mhz = (clocks.get("sm_max_mhz") or 1965)
        peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        achieved = bytes_per_event * events_per_launch / (k1_avg_ms * 1e-3) / 1e9
        ncu_file = os.path.join(ROOT, "profiles", "k1_ncu_summary.json")
        ncu_k1 = json.load(open(ncu_file)) if os.path.exists(ncu_file) else {}
        smem_traffic = ncu_k1.get("smem_bytes_per_event_at_128B_per_wavefront")
        roofline = dict(bound="smem", kernel="aps::k1_lean_kernel<true,21,1056>", achieved=achieved, peak=peak, unit="GB/s",
                        frac=achieved / peak,
                        traffic=(smem_traffic * events_per_launch if smem_traffic else None),
                        traffic_note="shared-memory wavefronts x 128 B per launch from the ncu capture (DRAM traffic of K1 is ~0.1 GB per launch)",
                        peak_source="computed 148 SM x 128 B/clk x clocks.max.sm (shared-memory bandwidth is not in "
                                    "MEASURED_PEAKS.json); HBM traffic of K1 is only the observation rows",
                        algorithmic_bytes_per_event=bytes_per_event, events_per_launch=events_per_launch,
                        kernel_ms=k1_avg_ms, kernel_share_of_step=k1_avg_ms * args.steps / total_ms,
                        ncu=(json.load(open(os.path.join(ROOT, "profiles", "k1_ncu_summary.json")))
                             if os.path.exists(os.path.join(ROOT, "profiles", "k1_ncu_summary.json")) else None))
        cpu_baseline = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_rep = min(256, 4 * cores)
            ev, dt = 0, 0.0
            for s in range(8):                       # bounded: ~10-20 s of CPU work
                ev_s, dt_s = oracle_sample(n_rep, 5.0, cores, seed0=s)
                ev, dt = ev + ev_s, dt + dt_s
                if dt > 12:
                    break
            cpu_baseline = dict(value=ev / dt, unit=UNIT, cores=cores, kind="port",
                                sample=f"{n_rep} replicas per pass (betas spread over the sweep), run(T=5), {cores} pthreads, "
                                       "C restatement of the reference algorithm; the reference's numpy path itself "
                                       "measured 3.4-4.1e3 events/s/core in the build container (BASELINE.md)")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic", config=dict(CONFIG, T=args.T, total_replicas=N_BETA * reps,
                                                                 mean_particles=mean_n, events_per_step_per_gpu=events_per_launch,
                                                                 guard_fallbacks=guard, replicas_not_done=bad),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=info["h2d_bytes"], d2h_bytes_per_step=info["d2h_bytes"],
                             api="launcher.sweep_over_betas (host parameters -> host reducers + profiles)"),
                    gpu_launches=int(launches), clocks=clocks, roofline=roofline, roofline_k2=k2,
                    cpu_baseline=cpu_baseline)
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
