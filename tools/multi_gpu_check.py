#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one process per GPU, NCCL):
  1. launcher: the sharded sweep (all_gather of reducers, all_reduce of profile sums) equals the same
     sweep computed unsharded on this rank alone;
  2. K2: the slab-decomposed lattice (ghost zones refreshed over NVLink P2P) is bit-identical to the
     single-GPU run of the same lattice."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import launcher as la
from aps_b200.sublattice import SublatticeLattice, TILE

rank, world = la.init_distributed_from_env()
if world < 2:
    print("multi_gpu_check: run under torchrun with at least 2 ranks (one per GPU); nothing to check with one rank")
    sys.exit(0)
PS = dict(L=200, xlim=1, rate_diffusion=0.1, rate_active=4, flip_rate_fn=None, init="poisson", N=110, scale_rates=False,
          local_kernel_sigma=0.02, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False)
RUN = dict(T=4.0, obs_dt=0.1)
g = la.make_exp_gradient(L=200, N=110, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
ik = dict(rho0_plus=g[0], rho0_minus=g[1])
betas, runs = np.linspace(0, 3, 7), 9
out = la.sweep_over_betas(betas, runs, PS, ik, RUN, base_seed=17)
spec = la.build_beta_sweep_spec(betas, runs, PS, ik, RUN, base_seed=17)
ens = la.DeviceEnsemble(spec, 0, len(spec.betas)).step()
scal = ens.pack_scalars().cpu().numpy()
assert np.array_equal(out["n_events"].ravel(), scal[:, 8].astype(np.int64)), "event counts differ"
red = scal[:, :8].reshape(len(betas), runs, 8)
np.testing.assert_array_equal(out["means"], red[:, :, 0].mean(1))
np.testing.assert_allclose(out["rho_plus_profile_mean"], (ens.prof.cpu().numpy()[:, 0] / runs), rtol=1e-13, atol=1e-15)
print(f"rank {rank}/{world}: launcher sharding OK, shard={out['info']['shard']}, events={int(out['n_events'].sum())}", flush=True)

kw = dict(D=0.3, lam=3.0, beta=1.2, dt=0.01, sigma_sites=3.0, seed=11)
lat = SublatticeLattice(8 * TILE, **kw)
lat.init_random(0.5, 0.6)
lat.refresh_every = 24
lat.run_passes(90)
full = lat.gather_state()
one = SublatticeLattice(8 * TILE, single_rank=True, **kw)
one.init_random(0.5, 0.6)
one.run_passes(90)
assert np.array_equal(full, one.state.cpu().numpy()), "slab run differs from single-GPU run"
rp, rm = lat.profile(32)
# global-magnetisation mode across ranks: own flips only + one 8-byte all-reduce per pass
kwg = dict(D=0.3, lam=3.0, beta=1.2, dt=0.01, sigma_sites=None, seed=13)
latg = SublatticeLattice(8 * TILE, **kwg)
latg.init_random(0.5, 0.6)
latg.refresh_every = 24
latg.run_passes(60)
fullg = latg.gather_state()
oneg = SublatticeLattice(8 * TILE, single_rank=True, **kwg)
oneg.init_random(0.5, 0.6)
oneg.run_passes(60)
assert np.array_equal(fullg, oneg.state.cpu().numpy()), "global-field slab run differs from single-GPU run"
assert int(latg.msum[0][0]) == int(oneg.msum[0][0])
print(f"rank {rank}/{world}: K2 global-field slabs bit-identical to single GPU, sum(sigma)={int(latg.msum[0][0])}", flush=True)
rp1, rm1 = one.profile(32)
assert np.array_equal(rp, rp1) and np.array_equal(rm, rm1)
print(f"rank {rank}/{world}: K2 slab decomposition bit-identical to single GPU (own sites {lat.own_lo}..{lat.own_hi})", flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
