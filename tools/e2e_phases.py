"""Host-side profile of the end-to-end call bench.py times (launcher.sweep_over_betas, config 2): cProfile over three calls."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from aps_b200 import launcher as la
w = B.WORKLOADS["config2"]
ik = B.init_kwargs("config2")
betas = np.linspace(0, 3, B.N_BETA)
def sync(): torch.cuda.synchronize()
la.sweep_over_betas(betas, B.REPS_PER_BETA, w["ps"], ik, w["run"], base_seed=8); sync()
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for i in range(3):
    la.sweep_over_betas(betas, B.REPS_PER_BETA, w["ps"], ik, w["run"], base_seed=9 + i)
sync(); print("sweep_over_betas ms per call", 1e3 * (time.perf_counter() - t0) / 3)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(18)
