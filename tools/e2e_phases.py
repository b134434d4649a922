import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from aps_b200 import launcher as la
ik = B.init_kwargs()
betas = np.linspace(0, 3, B.N_BETA)
run_kwargs = dict(B.RUN_KWARGS, T=20.0)
def sync(): torch.cuda.synchronize()
for it in range(3):
    t=[time.perf_counter()]
    spec = la.build_beta_sweep_spec(betas, B.REPS_PER_BETA, B.PS_KWARGS, ik, run_kwargs, base_seed=100+it); t.append(time.perf_counter())
    order = la.schedule_order(spec, 1); ps = la.permute_spec(spec, order); t.append(time.perf_counter())
    ens = la.DeviceEnsemble(ps, 0, len(spec.betas)); sync(); t.append(time.perf_counter())
    ens.init_particles(); sync(); t.append(time.perf_counter())
    ens.rb.run_philox(); sync(); t.append(time.perf_counter())
    ens.red = ens.rb.reduce(); sync(); t.append(time.perf_counter())
    per_rep = ens.rb.profile_sums(1); sync(); t.append(time.perf_counter())
    prof = torch.zeros((ens.n_points, 4, 1000), dtype=torch.float64, device=ens.dev).index_add_(0, ens.point_local, per_rep); sync(); t.append(time.perf_counter())
    scal = ens.pack_scalars().cpu().numpy(); ph = prof.cpu().numpy(); t.append(time.perf_counter())
    del ens; sync(); t.append(time.perf_counter())
    names=["spec","order","ensemble alloc+h2d","init","K1","reduce","profile_sums","index_add","d2h","free"]
    print(it, {n: round(1e3*(b-a),2) for n,a,b in zip(names,t[:-1],t[1:])})
t0=time.perf_counter(); out=la.sweep_over_betas(betas, B.REPS_PER_BETA, B.PS_KWARGS, ik, run_kwargs, base_seed=7); sync(); print("sweep_over_betas total ms", 1e3*(time.perf_counter()-t0))
import cProfile, pstats
run_kwargs = dict(B.RUN_KWARGS, T=20.0)
la.sweep_over_betas(betas, B.REPS_PER_BETA, B.PS_KWARGS, ik, run_kwargs, base_seed=8); sync()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for i in range(3):
    la.sweep_over_betas(betas, B.REPS_PER_BETA, B.PS_KWARGS, ik, run_kwargs, base_seed=9 + i)
sync(); print("T=20 sweep_over_betas ms per call", 1e3 * (time.perf_counter() - t0) / 3)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(14)
