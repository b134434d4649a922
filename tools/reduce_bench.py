import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from aps_b200 import launcher as la, capi
lib = capi.load()
ik = B.init_kwargs(); betas = np.linspace(0, 3, B.N_BETA)
spec = la.build_beta_sweep_spec(betas, B.REPS_PER_BETA, B.PS_KWARGS, ik, dict(B.RUN_KWARGS, T=20.0), base_seed=1)
ens = la.DeviceEnsemble(spec, 0, len(spec.betas)); ens.init_particles(); ens.rb.run_philox(); torch.cuda.synchronize()
for nt in (64, 128, 256, 512):
    lib.aps_debug_set_reduce_threads(nt)
    ens.rb.reduce(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r = ens.rb.reduce()
    e1.record(); torch.cuda.synchronize()
    print(nt, "reduce ms", e0.elapsed_time(e1) / 5, float(r[:, 0].sum()))
e0.record()
for _ in range(5): ens.rb.profile_sums(1)
e1.record(); torch.cuda.synchronize(); print("profile_sums ms", e0.elapsed_time(e1) / 5)
