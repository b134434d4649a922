"""CPU stand-in for launcher.DeviceEnsemble (TEST HELPER): the same shard interface, computed with the
oracle (init + Philox run) and the numpy reducer port.  Used (a) under gloo to test sharding and
collectives without a GPU, (b) as the checker of the GPU launcher's end-to-end results."""
import numpy as np
import torch

from aps_b200 import capi
from aps_b200.capi import ApsInitArgs
from aps_b200.launcher import _model_params, expected_poisson_particles
from common import HostRun
from aps_b200.batch import make_params
from oracle import oracle, reducers_np as rn


class OracleEnsemble:
    def __init__(self, spec, lo, hi, device=None):
        self.spec, self.lo, self.hi = spec, lo, hi
        self.mp = mp = _model_params(spec.ps_kwargs)
        self.R = hi - lo
        T, obs_dt = float(spec.run_kwargs["T"]), float(spec.run_kwargs["obs_dt"])
        self.times_obs = np.arange(0.0, T, obs_dt)
        self.T = T
        if mp["init"] == "poisson":
            self.n_max = max(expected_poisson_particles(spec.profiles_plus[p], spec.profiles_minus[p], mp["K"])[1]
                             for p in range(len(spec.profiles_plus)))
        else:
            self.n_max = int(spec.N_of.max()) if spec.N_of is not None else mp["N"]
        self.n_max = max(8, (self.n_max + 7) // 8 * 8)
        self.h2d_bytes = 0
        self.n_points = int(spec.point_of.max()) + 1
        self.prof = None

        class _RB:
            M = len(self.times_obs)
        self.rb = _RB()

    def init_states(self):
        mp, spec, R = self.mp, self.spec, self.R
        sl = slice(self.lo, self.hi)
        seeds = np.ascontiguousarray(spec.seeds[sl], dtype=np.uint64)
        pos0 = np.zeros((R, self.n_max), np.int32); sg0 = np.ones((R, self.n_max), np.int8); n = np.zeros(R, np.int32)
        rp = np.ascontiguousarray(spec.profiles_plus, dtype=np.float64) if spec.profiles_plus is not None else None
        rm = np.ascontiguousarray(spec.profiles_minus, dtype=np.float64) if spec.profiles_minus is not None else None
        pof = np.ascontiguousarray(spec.profile_of[sl], dtype=np.int32) if spec.profile_of is not None else None
        nof = np.ascontiguousarray(spec.N_of[sl], dtype=np.int32) if spec.N_of is not None else None
        a = ApsInitArgs(R, mp["L"], mp["K"], self.n_max, 1 if mp["init"] == "poisson" else 0, mp["N"],
                        len(rp) if rp is not None else 0, 0,
                        rp.ctypes.data if rp is not None else None, rm.ctypes.data if rm is not None else None,
                        pof.ctypes.data if pof is not None else None, nof.ctypes.data if nof is not None else None,
                        seeds.ctypes.data, pos0.ctypes.data, sg0.ctypes.data, n.ctypes.data)
        assert oracle.load().aps_oracle_init(a) == 0
        return seeds, pos0, sg0, n

    def step(self, want_profiles=True, threads=8):
        mp, spec = self.mp, self.spec
        seeds, pos0, sg0, n = self.init_states()
        sl = slice(self.lo, self.hi)
        hr = HostRun(mp["L"], self.n_max, len(self.times_obs), n, pos0, sg0, spec.betas[sl], self.times_obs,
                     mp["weights"], seeds=seeds, record=3)
        P = make_params(mp["L"], mp["K"], mp["radius"], mp["D"], mp["lam"], self.T,
                        (capi.APS_FLAG_CROWDING if mp["crowding"] else 0) | (capi.APS_FLAG_PERIODIC if mp.get("periodic") else 0))
        assert oracle.load().aps_oracle_run(P, hr.batch, 1, threads) == 0
        self.hr = hr
        red = np.zeros((self.R, capi.APS_RED_N))
        L, dx, M = mp["L"], mp["dx"], len(self.times_obs)
        prof = np.zeros((self.n_points, 4, L))
        mbar = np.zeros(self.R); hist = np.zeros((max(1, self.n_points), 256), np.int64)
        for r in range(self.R):
            nobs, nn = int(hr.n_obs[r]), int(n[r])
            denom = float(max(1, nn)) * dx
            rho_p = np.zeros((M, L)); rho_m = np.zeros((M, L))
            rho_p[:nobs] = hr.obs_cp[r, :nobs].astype(np.int64) / denom
            rho_m[:nobs] = hr.obs_cm[r, :nobs].astype(np.int64) / denom
            total = rho_p + rho_m
            mg = np.zeros(M); mg[:nobs] = hr.obs_sigma_sum[r, :nobs] / float(nn)
            pos_list = [hr.obs_pos[r, k, :nn].astype(np.int64) for k in range(M)]
            mean_v, v_eff, si, ei, frac = rn.v_eff_and_window(self.times_obs, total, L)
            d_eff = rn.d_eff_active(self.times_obs, pos_list, dx, si, ei) if nobs >= ei else np.nan
            red[r] = [mean_v, d_eff, rn.mean_magnetisation(mg, si, ei), rn.rho_eff(total, si, ei),
                      rn.blocking_probability(total, rho_p, si, ei), si, ei, nobs]
            mbar[r] = mg[M // 2:nobs].sum() / max(1, nobs - M // 2)        # aps_m_histogram_device: rows [M/2, n_obs)
            b = int(np.floor((mbar[r] + 1.0) / 2.0 * 256))
            hist[int(spec.point_of[self.lo + r]), min(255, max(0, b))] += 1
            if want_profiles:
                lo_r, hi_r = M // 2, M
                g = int(spec.point_of[self.lo + r])
                mp_, mm_ = rho_p[lo_r:hi_r].mean(0), rho_m[lo_r:hi_r].mean(0)
                prof[g, 0] += mp_; prof[g, 1] += mm_; prof[g, 2] += mp_ ** 2; prof[g, 3] += mm_ ** 2
        self.red = red
        self.mbar, self.hist = mbar, torch.from_numpy(hist)
        self.n = n
        self.prof = torch.from_numpy(prof) if want_profiles else None
        return self

    def pack_scalars(self):
        hr = self.hr
        return torch.from_numpy(np.concatenate([self.red, hr.n_events[:, None].astype(float),
                                                hr.status[:, None].astype(float), self.n[:, None].astype(float),
                                                self.mbar[:, None]], axis=1))
