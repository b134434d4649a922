"""Shared helpers for the parity tests: fixture loading, oracle driver, comparison."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from aps_b200 import capi
from aps_b200.batch import make_batch, make_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# Named custom flip rates of the `custom_flip_*` fixtures (callables cannot be stored in a fixture): the SAME Python
# callables are handed to the reference (tools/gen_golden.py) and to the drop-in / tabulated for the oracle.
FLIP_FNS = {
    "glauber_1p5": lambda sigma, m: 1.0 - sigma * np.tanh(1.5 * m),
    "exp_quadratic": lambda sigma, m: np.exp(-0.8 * sigma * m) + 0.25 * m * m,
}


def case_names():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if not n.startswith(("stat_", "pde_", "full_"))]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def params_from_case(c, T=None):
    m = c["meta"]
    flags = capi.APS_FLAG_CROWDING if m["ps"].get("crowding_suppresses_rates") else 0
    flags |= capi.APS_FLAG_PERIODIC if m.get("periodic") else 0
    if m.get("anchors"):
        flags |= (capi.APS_FLAG_SUPPRESS_FLIP_BOUND if m["suppress"] else 0) | (capi.APS_FLAG_IMMOBILIZE if m["immobilize"] else 0)
    return make_params(m["L"], m["K"], m["radius"], m["rate_diffusion"], m["rate_active"],
                       m["run"]["T"] if T is None else T, flags, k_on=m.get("k_on", 0.0), k_off=m.get("k_off", 0.0),
                       k_exit=m.get("k_exit", 0.0))


class HostRun:
    """Host-side buffers for one batch; `.batch` is the aps_batch pointing at them."""

    def __init__(self, L, n_max, M, n, pos0, sigma0, beta, times_obs, weights, draws=None, draw_off=None,
                 seeds=None, record=7, trace_cap=0, max_events=0, t_start=None, obs_start=None, ev_start=None,
                 anchor_mask=None, bound0=None, exit_cap=0, alloc_m_local=True, flip_tab=None):
        R = len(n)
        self.R, self.L, self.n_max, self.M = R, L, n_max, M
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.pos0 = np.ascontiguousarray(pos0, dtype=np.int32).reshape(R, n_max)
        self.sigma0 = np.ascontiguousarray(sigma0, dtype=np.int8).reshape(R, n_max)
        self.beta = np.ascontiguousarray(beta, dtype=np.float64)
        self.times_obs = np.ascontiguousarray(times_obs, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64) if weights is not None and len(weights) else None
        self.draws = None if draws is None else np.ascontiguousarray(draws, dtype=np.float64)
        self.draw_off = None if draw_off is None else np.ascontiguousarray(draw_off, dtype=np.int64)
        self.seeds = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        self.t_start = None if t_start is None else np.ascontiguousarray(t_start, dtype=np.float64)
        self.obs_start = None if obs_start is None else np.ascontiguousarray(obs_start, dtype=np.int32)
        self.ev_start = None if ev_start is None else np.ascontiguousarray(ev_start, dtype=np.int64)
        Mr = max(M, 1)
        self.obs_cp = np.full((R, Mr, L), -7, np.int8)
        self.obs_cm = np.full((R, Mr, L), -7, np.int8)
        self.obs_pos = np.full((R, Mr, n_max), -1, np.int32)
        self.obs_sigma_sum = np.full((R, Mr), -99999, np.int32)
        self.obs_m_local = np.full((R, Mr, L), np.nan, np.float64) if alloc_m_local else None   # 8 B/site/row: skip for big batches
        self.n_obs = np.zeros(R, np.int32)
        self.n_events = np.zeros(R, np.int64)
        self.t_end = np.zeros(R, np.float64)
        self.status = np.full(R, -1, np.int32)
        self.n_guard = np.zeros(R, np.int64)
        self.draws_used = np.zeros(R, np.int64)
        self.pos_end = np.full((R, n_max), -1, np.int32)
        self.sigma_end = np.zeros((R, n_max), np.int8)
        self.trace = np.full((R, max(trace_cap, 1), 3), -5, np.int32) if trace_cap else None
        self.anchor_mask = None if anchor_mask is None else np.ascontiguousarray(anchor_mask, dtype=np.uint8)
        self.bound0 = None if bound0 is None else np.ascontiguousarray(bound0, dtype=np.int8).reshape(R, n_max)
        self.n_end = np.full(R, -1, np.int32)
        self.bound_end = np.zeros((R, n_max), np.int8)
        self.obs_n = np.full((R, Mr), -1, np.int32)
        self.obs_bound = np.zeros((R, Mr, n_max), np.int8)
        self.exit_t = np.full((R, max(exit_cap, 1)), np.nan, np.float64)
        self.exit_pos = np.full((R, max(exit_cap, 1)), -1, np.int32)
        self.n_exit = np.zeros(R, np.int32)
        self.flip_tab = None if flip_tab is None else np.ascontiguousarray(flip_tab, dtype=np.float64).reshape(2, -1)
        self.batch, self._keep = make_batch(
            R, n_max, M, record=record, max_events=max_events, trace_cap=trace_cap, exit_cap=exit_cap,
            flip_tab=self.flip_tab, flip_G=0 if flip_tab is None else self.flip_tab.shape[1] - 1,
            anchor_mask=self.anchor_mask, bound0=self.bound0, n_end=self.n_end, bound_end=self.bound_end, obs_n=self.obs_n,
            obs_bound=self.obs_bound, exit_t=self.exit_t, exit_pos=self.exit_pos, n_exit=self.n_exit,
            times_obs=self.times_obs, weights=self.weights, beta=self.beta, n=self.n, pos0=self.pos0,
            sigma0=self.sigma0, draws=self.draws, draw_off=self.draw_off, seeds=self.seeds,
            t_start=self.t_start, obs_start=self.obs_start, ev_start=self.ev_start,
            obs_cp=self.obs_cp, obs_cm=self.obs_cm, obs_pos=self.obs_pos, obs_sigma_sum=self.obs_sigma_sum,
            obs_m_local=self.obs_m_local, n_obs=self.n_obs, n_events=self.n_events, t_end=self.t_end,
            status=self.status, n_guard=self.n_guard, draws_used=self.draws_used, pos_end=self.pos_end,
            sigma_end=self.sigma_end, trace=self.trace)

    OUT_FIELDS = ["obs_cp", "obs_cm", "obs_pos", "obs_sigma_sum", "obs_m_local", "n_obs", "n_events", "t_end",
                  "status", "draws_used", "pos_end", "sigma_end", "trace", "n_end", "bound_end", "obs_n", "obs_bound",
                  "exit_t", "exit_pos", "n_exit"]


def hostrun_from_case(c, trace=True, **kw):
    m = c["meta"]
    n = m["n"]
    if m.get("flip"):
        from aps_b200.engine import tabulate_flip_rate
        kw.setdefault("flip_tab", tabulate_flip_rate(FLIP_FNS[m["flip"]]))
    return HostRun(m["L"], max(n, 1), len(c["times_obs"]), [n], c["pos0"], c["sigma0"], [m["ps"]["beta"]],
                   c["times_obs"], c["weights"], draws=c["draws"], draw_off=[0, len(c["draws"])],
                   trace_cap=(len(c["trace"]) + 4) if trace else 0,
                   anchor_mask=c["anchor_mask"] if m.get("anchors") else None,
                   exit_cap=(len(c["exit_times"]) + 4) if m.get("anchors") else 0, **kw)


def run_oracle(params, hr, mode=0, threads=1):
    from oracle import oracle

    rc = oracle.load().aps_oracle_run(params, hr.batch, mode, threads)
    assert rc == 0
    return hr


def assert_same_outputs(a: HostRun, b: HostRun, bitwise_time=True):
    """GPU vs oracle: every output field identical (m_local and t_end compared as bit patterns)."""
    for f in HostRun.OUT_FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        if x is None and y is None:
            continue
        if f == "t_end" and not bitwise_time:
            np.testing.assert_allclose(x, y, rtol=1e-12)
            continue
        if x.dtype == np.float64:
            assert np.array_equal(x.view(np.uint64), y.view(np.uint64)), f"field {f} differs"
        else:
            assert np.array_equal(x, y), f"field {f} differs"


def assert_matches_reference(c, hr: HostRun, rep=0):
    """Outputs of a run (oracle or GPU) against what the unmodified reference returned."""
    m = c["meta"]
    n, L, n_obs, n_ev = m["n"], m["L"], m["n_obs"], m["n_events"]
    assert hr.status[rep] == capi.APS_RUN_DONE
    assert hr.n_obs[rep] == n_obs
    assert hr.n_events[rep] == n_ev
    assert hr.draws_used[rep] == len(c["draws"])
    if hr.trace is not None:
        assert np.array_equal(hr.trace[rep, :n_ev], c["trace"]), "event trace differs from the reference"
    n_t = hr.obs_n[rep, :n_obs].astype(np.int64)                  # particle count per row (exits shrink it)
    if "count_obs" in c:
        assert np.array_equal(n_t, c["count_obs"][:n_obs])
    else:
        assert (n_t == n).all()
    denom = np.maximum(1, n_t).astype(float)[:, None] * m["dx"]
    rho_p = hr.obs_cp[rep, :n_obs].astype(np.int64) / denom      # CLASS.py:205-213
    rho_m = hr.obs_cm[rep, :n_obs].astype(np.int64) / denom
    assert np.array_equal(rho_p, c["rho_p_list"][:n_obs])
    assert np.array_equal(rho_m, c["rho_m_list"][:n_obs])
    assert np.array_equal(rho_p + rho_m, c["total_list"][:n_obs])
    for row in range(n_obs):
        k = int(n_t[row])
        assert np.array_equal(hr.obs_pos[rep, row, :k], c["pos_obs"][row, :k])
        if "bound_obs" in c:
            assert np.array_equal(hr.obs_bound[rep, row, :k], c["bound_obs"][row, :k])
    assert np.array_equal(hr.obs_sigma_sum[rep, :n_obs] / n_t.astype(float), c["m_global"][:n_obs])
    if "exit_times" in c:
        ne = len(c["exit_times"])
        assert hr.n_exit[rep] == ne and hr.n_end[rep] == n - ne
        assert np.array_equal(hr.exit_pos[rep, :ne], c["exit_positions"])
        np.testing.assert_allclose(hr.exit_t[rep, :ne], c["exit_times"], rtol=1e-13)   # clock: exp/log differ by <= 1 ulp
    got, want = hr.obs_m_local[rep, :n_obs], c["m_local_list"][:n_obs]
    if m.get("periodic") and m["radius"] >= 0:
        # the reference convolves by FFT (CLASS.py:224-227); the direct ring sum agrees to rounding, not bitwise
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
        return
    assert np.array_equal(got, want), f"m_local differs: max abs {np.abs(got - want).max()}"
    # rows never reached stay zero in the reference (CLASS.py:466-472)
    assert not c["rho_p_list"][n_obs:].any()


def _sha(a):
    import hashlib

    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a + 0.0                      # -0.0 and +0.0 compare equal in the reference's own tests of equality
    return hashlib.sha256(a.tobytes()).hexdigest()[:32]


def digest_out(out):
    """Digests of a `ParticleSystem.run()` dict (reference or drop-in): the full-length pins of
    tests/golden/full_length.json (tools/gen_golden_full.py).  FFT arrays are post-processing of total_list by a
    different FFT library on the device and are compared by value elsewhere, not digested."""
    n_obs = sum(p is not None for p in out["pos_list"])
    d = dict(n_obs=int(n_obs),
             pos=_sha(np.stack([np.asarray(p, dtype=np.int64) for p in out["pos_list"][:n_obs]])),
             counts=[int(x) for x in out["particle_count_list"][:n_obs]][:4])
    for k in ["rho_p_list", "rho_m_list", "total_list", "m_local_list", "m_global"]:
        d[k] = _sha(out[k])
    if out.get("var_list") is not None:
        d["var_list"] = _sha(out["var_list"])
    return d
