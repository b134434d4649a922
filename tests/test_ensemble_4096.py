"""BASELINE target: "a beta-sweep ensemble of at least 4096 replicas runs bit-exact in replay mode".

4096 replicas (64 beta x 64 runs) of the sweep_beta configuration with run(T=2):
  1. native run with the event trace recorded;
  2. the consumed Philox variates are written out as replay logs (oracle helper) and the whole ensemble is
     re-run in REPLAY mode: every output (state, observation rows, clock, event counts) must be bit-identical;
  3. a 64-replica subset (one per beta) is checked bit-for-bit against the CPU oracle replaying the same logs.
Size-independent properties on all 4096: particle conservation, exclusion, single-file order (K = 1 nearest-
neighbour hops never reorder particles), sigma-sum consistency."""
import numpy as np
import pytest
import torch

from aps_b200 import capi, launcher as la
from aps_b200.capi import APS_REC_COUNTS, APS_REC_POS
from aps_b200.engine import ReplicaBatch

pytestmark = pytest.mark.gpu

PS = dict(L=1000, xlim=1, rate_diffusion=0.02, rate_active=5, flip_rate_fn=None, init="poisson", N=500, scale_rates=False,
          local_kernel_sigma=0.005, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False)


def test_4096_replicas_replay_equals_native_and_oracle():
    from common import HostRun, assert_same_outputs, run_oracle
    from aps_b200.batch import make_params
    from oracle import oracle
    g = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
    g2 = la.make_exp_gradient(L=1000, N=500, frac_plus=0.75, decay_length=0.2, anchor_positions=None)
    T, obs_dt, cap = 2.0, 0.1, 2600
    spec = la.build_beta_sweep_spec(np.linspace(0, 3, 64), 64, PS, dict(rho0_plus=g[0], rho0_minus=g2[1]),
                                    dict(T=T, obs_dt=obs_dt), base_seed=3)
    ens = la.DeviceEnsemble(spec, 0, 4096)
    ens.init_particles()
    rb = ens.rb
    trace = torch.full((4096, cap, 3), -5, dtype=torch.int32, device="cuda")
    b, keep = rb._batch(seeds=rb.seeds, trace=trace)
    b.trace_cap = cap
    capi.check(rb.lib.aps_run_philox_device(rb.params, b, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    nat = {k: getattr(rb, k).clone() for k in ["obs_cp", "obs_cm", "obs_pos", "obs_sigma_sum", "n_obs", "n_events", "t_end",
                                                "status", "pos_end", "sigma_end"]}
    nev = nat["n_events"].cpu().numpy()
    assert (nat["status"] == 0).all() and nev.max() < cap and nev.min() > 500
    # ---- size-independent properties on the whole ensemble ----
    n = ens.n.cpu().numpy()
    tot = (nat["obs_cp"].to(torch.int32) + nat["obs_cm"].to(torch.int32))
    assert int(tot.max()) <= 1                                                   # exclusion, K = 1
    assert np.array_equal(tot.sum(dim=2).cpu().numpy(), np.repeat(n[:, None], tot.shape[1], 1))   # conservation
    s_sum = (nat["obs_cp"].to(torch.int32) - nat["obs_cm"].to(torch.int32)).sum(dim=2)
    assert torch.equal(s_sum, nat["obs_sigma_sum"])
    pos = nat["obs_pos"].cpu().numpy()
    for r in range(0, 4096, 97):                                                 # single-file order is preserved
        d = np.diff(pos[r, :, :n[r]], axis=1)
        assert (d > 0).all()
    # ---- native variates -> replay logs ----
    kinds = trace[:, :, 1].cpu().numpy()
    seeds = np.asarray(spec.seeds, dtype=np.uint64)
    lib = oracle.load()
    logs, off = [], [0]
    buf = np.zeros(4 * cap)
    for r in range(4096):
        k = np.ascontiguousarray(kinds[r, :nev[r]], dtype=np.int32)
        w = lib.aps_oracle_philox_log(int(seeds[r]), 0, int(nev[r]), k.ctypes.data, buf.ctypes.data)
        logs.append(buf[:w].copy()); off.append(off[-1] + w)
    draws = torch.from_numpy(np.concatenate(logs)).cuda()
    draw_off = torch.tensor(off, dtype=torch.int64, device="cuda")
    for k in ["obs_cp", "obs_cm", "obs_pos", "obs_sigma_sum", "n_obs", "n_events", "t_end", "status", "pos_end", "sigma_end"]:
        getattr(rb, k).zero_()
    rb.run_replay(draws, draw_off)
    torch.cuda.synchronize()
    for k, v in nat.items():
        got = getattr(rb, k)
        assert torch.equal(got.view(torch.int64) if got.dtype == torch.float64 else got,
                           v.view(torch.int64) if v.dtype == torch.float64 else v), f"replay differs from native in {k}"
    assert torch.equal(rb.draws_used, draw_off[1:] - draw_off[:-1]) and int(rb.n_guard.sum()) == 0
    # ---- one replica per beta against the CPU oracle (replay of the same logs) ----
    idx = np.arange(0, 4096, 64)
    mp = ens.mp
    pos0 = ens.pos0.cpu().numpy()[idx]; sg0 = ens.sigma0.cpu().numpy()[idx]
    sub_logs = [logs[i] for i in idx]
    sub_off = np.concatenate([[0], np.cumsum([len(x) for x in sub_logs])])
    hr = HostRun(mp["L"], ens.n_max, rb.M, n[idx], pos0, sg0, np.asarray(spec.betas)[idx], ens.times_obs, mp["weights"],
                 draws=np.concatenate(sub_logs), draw_off=sub_off, record=3)
    run_oracle(make_params(mp["L"], 1, mp["radius"], mp["D"], mp["lam"], T), hr, threads=8)
    assert np.array_equal(hr.obs_cp, nat["obs_cp"].cpu().numpy()[idx])
    nat_pos = nat["obs_pos"].cpu().numpy()[idx]
    for j, r in enumerate(idx):                    # slots beyond n[r] are padding
        assert np.array_equal(hr.obs_pos[j, :, :n[r]], nat_pos[j, :, :n[r]])
    assert np.array_equal(hr.n_events, nev[idx]) and np.array_equal(hr.t_end.view(np.uint64), nat["t_end"].cpu().numpy()[idx].view(np.uint64))
    nat_end = nat["pos_end"].cpu().numpy()[idx]
    assert all(np.array_equal(hr.pos_end[j, :n[r]], nat_end[j, :n[r]]) for j, r in enumerate(idx)) and (hr.status == 0).all()
