"""BASELINE target: "a beta-sweep ensemble of at least 4096 replicas runs bit-exact in replay mode".

All 4096 replicas (64 beta x 64 runs) of BASELINE config 2 at its FULL length, run(T=20, obs_dt=0.1): ~5.9e7 events.
  1. the CPU oracle runs every replica (native Philox streams, event trace recorded) in chunks of 256;
  2. the replay logs are written from the ORACLE's trajectories (`aps_oracle_philox_log`: the variates in the order
     the reference draws them, the 4th one only for the oracle's diffusive events) — the kernel has no part in them;
  3. the GPU replays all 4096 logs in one launch: every output of every replica (observation rows, positions,
     sigma sums, event counts, final state) must equal the oracle's; the event clock agrees to 1e-12 here, because
     replay mode sums R in numpy's pairwise order (the reference's clock) and native mode takes the selection
     scan's total (aps_math.h, aps_native_total) — 64 replicas are additionally replayed by the ORACLE, and there
     the GPU's clock must match bit for bit;
  4. the GPU's native Philox run must equal the oracle's native run in every output INCLUDING the bit pattern of
     the event clock (same definition of R on both sides).
Size-independent properties on all 4096: particle conservation, exclusion, single-file order (K = 1 nearest-
neighbour hops never reorder particles), sigma-sum consistency.
The oracle needs ~2 minutes on 16 host threads for step 1; set APS_TEST_ENSEMBLE_T to shorten the run."""
import os

import numpy as np
import pytest
import torch

from aps_b200 import launcher as la

pytestmark = pytest.mark.gpu

PS = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, flip_rate_fn=None, init="poisson", N=500, scale_rates=False,
          local_kernel_sigma=0.005, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False)
OUT_KEYS = ["obs_cp", "obs_cm", "obs_pos", "obs_sigma_sum", "n_obs", "n_events", "t_end", "status", "pos_end", "sigma_end"]


def _bits(t):
    return t.view(torch.int64) if t.dtype == torch.float64 else t


def test_4096_replicas_full_length_replay_and_native_equal_the_oracle():
    from common import HostRun, run_oracle
    from aps_b200.batch import make_params
    from oracle import oracle
    T = float(os.environ.get("APS_TEST_ENSEMBLE_T", "20.0"))
    obs_dt, R, CH = 0.1, 4096, 256
    g = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
    g2 = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.2, anchor_positions=None)
    spec = la.build_beta_sweep_spec(np.linspace(0, 3, 64), 64, PS, dict(rho0_plus=g[0], rho0_minus=g2[1]),
                                    dict(T=T, obs_dt=obs_dt), base_seed=3)
    ens = la.DeviceEnsemble(spec, 0, R)
    ens.init_particles()                      # device Philox init (bit-exact against the oracle: test_k1_parity)
    rb, mp = ens.rb, ens.mp
    n = ens.n.cpu().numpy()
    pos0, sg0 = ens.pos0.cpu().numpy(), ens.sigma0.cpu().numpy()
    seeds = np.asarray(spec.seeds, dtype=np.uint64)
    betas = np.asarray(spec.betas, dtype=np.float64)
    params = make_params(mp["L"], 1, mp["radius"], mp["D"], mp["lam"], T)
    lib = oracle.load()
    threads = os.cpu_count() or 8
    cap = int(1150 * T) + 400                 # events per replica: ~750-850 per unit time

    # ---- 1+2: oracle trajectories (native streams) and the replay logs they imply ----
    ora = {k: [] for k in OUT_KEYS}
    logs, off = [], [0]
    buf = np.zeros(4 * cap)
    for lo in range(0, R, CH):
        sl = slice(lo, lo + CH)
        hr = HostRun(mp["L"], ens.n_max, rb.M, n[sl], pos0[sl], sg0[sl], betas[sl], ens.times_obs, mp["weights"],
                     seeds=seeds[sl], record=3, trace_cap=cap, alloc_m_local=False)
        run_oracle(params, hr, mode=1, threads=threads)
        assert (hr.status == 0).all() and hr.n_events.max() < cap
        for k in OUT_KEYS:
            ora[k].append(getattr(hr, k).copy())
        for j in range(CH):
            kinds = np.ascontiguousarray(hr.trace[j, :hr.n_events[j], 1], dtype=np.int32)
            w = lib.aps_oracle_philox_log(int(seeds[lo + j]), 0, int(hr.n_events[j]), kinds.ctypes.data, buf.ctypes.data)
            logs.append(buf[:w].copy()); off.append(off[-1] + w)
        del hr
    ora = {k: np.concatenate(v) for k, v in ora.items()}
    nev = ora["n_events"]
    assert nev.sum() > 2.5e6 * T and nev.min() > 500 * T

    def assert_equals_oracle(tag, clock_bitwise):
        torch.cuda.synchronize()
        for k in OUT_KEYS:
            got = getattr(rb, k).cpu().numpy()
            want = ora[k]
            if k == "t_end" and not clock_bitwise:
                np.testing.assert_allclose(got, want, rtol=1e-12, err_msg=f"{tag}: event clock")
                continue
            if k in ("obs_pos", "pos_end", "sigma_end"):          # slots beyond n[r] are padding
                mask = np.arange(ens.n_max)[None, :] < n[:, None]
                mask = mask[:, None, :] if k == "obs_pos" else mask
                got, want = np.where(mask, got, 0), np.where(mask, want, 0)
            if got.dtype == np.float64:
                got, want = got.view(np.uint64), want.view(np.uint64)
            bad = np.nonzero((got != want).reshape(R, -1).any(axis=1))[0]
            assert bad.size == 0, f"{tag}: {k} differs from the oracle for {bad.size} replicas (first {bad[:5]})"

    # ---- 3: GPU replay of the oracle's logs ----
    draws = torch.from_numpy(np.concatenate(logs)).cuda()
    draw_off = torch.tensor(off, dtype=torch.int64, device="cuda")
    del logs
    rb.run_replay(draws, draw_off)
    assert_equals_oracle("replay", clock_bitwise=False)
    assert torch.equal(rb.draws_used, draw_off[1:] - draw_off[:-1])
    rep = {k: getattr(rb, k).clone() for k in OUT_KEYS}
    # replay-mode clock, bit for bit: the oracle replays the same logs for one replica per beta
    idx = np.arange(0, R, 64)
    d_h, off_h = draws.cpu().numpy(), np.asarray(off)
    sub = [d_h[off_h[i]:off_h[i + 1]] for i in idx]
    hr = HostRun(mp["L"], ens.n_max, rb.M, n[idx], pos0[idx], sg0[idx], betas[idx], ens.times_obs, mp["weights"],
                 draws=np.concatenate(sub), draw_off=np.concatenate([[0], np.cumsum([len(x) for x in sub])]), record=3,
                 alloc_m_local=False)
    run_oracle(params, hr, mode=0, threads=threads)
    assert np.array_equal(hr.t_end.view(np.uint64), rep["t_end"].cpu().numpy()[idx].view(np.uint64))
    assert np.array_equal(hr.n_events, nev[idx]) and np.array_equal(hr.obs_cp, rep["obs_cp"].cpu().numpy()[idx])
    del draws, d_h

    # ---- 4: GPU native run == oracle native run, clock included ----
    for k in OUT_KEYS:
        getattr(rb, k).zero_()
    rb.run_philox()
    assert_equals_oracle("native", clock_bitwise=True)

    # ---- size-independent properties on the whole ensemble ----
    tot = rep["obs_cp"].to(torch.int32) + rep["obs_cm"].to(torch.int32)
    assert int(tot.max()) <= 1                                                   # exclusion, K = 1
    assert np.array_equal(tot.sum(dim=2).cpu().numpy(), np.repeat(n[:, None], tot.shape[1], 1))   # conservation
    s_sum = (rep["obs_cp"].to(torch.int32) - rep["obs_cm"].to(torch.int32)).sum(dim=2)
    assert torch.equal(s_sum, rep["obs_sigma_sum"])
    pos = rep["obs_pos"].cpu().numpy()
    for r in range(0, R, 97):                                                    # single-file order is preserved
        assert (np.diff(pos[r, :, :n[r]], axis=1) > 0).all()
