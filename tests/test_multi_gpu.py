"""Multi-GPU launcher parity (one process per GPU, NCCL), run by `pytest -m gpu` on a box with >= 2 GPUs and skipped
otherwise: the sharded sweep (strided shards, all_gather of the per-run reducers, all_reduce of the per-beta profile sums
and of the device-side magnetisation histogram) equals the same sweep computed unsharded on one rank.
The K2 slab decomposition with its in-kernel NVLink exchange is tested in tests/test_k2.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))

PS = dict(L=200, xlim=1, rate_diffusion=0.1, rate_active=4, flip_rate_fn=None, init="poisson", N=110, scale_rates=False,
          local_kernel_sigma=0.02, periodic=False, anchor_positions=None, site_capacity=1, crowding_suppresses_rates=False)
RUN = dict(T=4.0, obs_dt=0.1)


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        from aps_b200 import launcher as la
        la.init_distributed_from_env()
        g = la.make_exp_gradient(L=200, N=110, frac_plus=0.75, decay_length=0.35, anchor_positions=None)
        ik = dict(rho0_plus=g[0], rho0_minus=g[1])
        betas, runs = np.linspace(0, 3, 7), 9
        out = la.sweep_over_betas(betas, runs, PS, ik, RUN, base_seed=17)
        spec = la.build_beta_sweep_spec(betas, runs, PS, ik, RUN, base_seed=17)
        ens = la.DeviceEnsemble(spec, 0, len(spec.betas)).step()          # the whole sweep on this rank alone
        scal = ens.pack_scalars().cpu().numpy()
        ok = np.array_equal(out["n_events"].ravel(), scal[:, 8].astype(np.int64))
        red = scal[:, :8].reshape(len(betas), runs, 8)
        ok = ok and np.array_equal(out["means"], red[:, :, 0].mean(1))
        ok = ok and np.allclose(out["rho_plus_profile_mean"], ens.prof.cpu().numpy()[:, 0] / runs, rtol=1e-13, atol=1e-15)
        ok = ok and np.array_equal(out["m_hist"], ens.hist.cpu().numpy()) and int(out["m_hist"].sum()) == len(betas) * runs
        ok = ok and np.array_equal(out["m_bar"].ravel(), ens.mbar.cpu().numpy())
        q.put((rank, bool(ok), str(out["info"]["shard"])))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_sweep_equals_unsharded(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 13 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = [q.get(timeout=600) for _ in range(world)]
    [p.join(120) for p in procs]
    for rank, ok, msg in got:
        assert ok, f"rank {rank}: sharded sweep differs from the unsharded one ({msg})"
