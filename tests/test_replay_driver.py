"""Host logic of the drop-in's replay driver (speculative chunks + bit-generator rewind,
particle_system.py:_run_replay) exercised on CPU: the device batch is replaced by a stand-in that
executes the same aps_batch descriptor with the oracle.  Checks that the chunked/rewound run
reproduces the reference's trajectory and leaves the Generator where the reference leaves it."""
import sys, os

import numpy as np
import pytest
import torch

from aps_b200.batch import make_batch
from common import HostRun, assert_matches_reference, load_case, params_from_case
from oracle import oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dropin"))


class _T:
    """minimal tensor-like view over a numpy array"""
    def __init__(self, a): self.a = a
    def item(self): return self.a.reshape(-1)[0].item()
    def copy_(self, o): self.a[...] = o.a


class OracleBackedBatch:
    def __init__(self, c, pos, sigma):
        m = c["meta"]
        n = m["n"]
        self.dev = "cpu"
        self.P = params_from_case(c)
        self.hr = HostRun(m["L"], n, len(c["times_obs"]), [n], pos.astype(np.int32), sigma, [m["ps"]["beta"]],
                          c["times_obs"], c["weights"], draws=np.zeros(4), draw_off=[0, 4],
                          anchor_mask=c["anchor_mask"] if m.get("anchors") else None, exit_cap=n)
        if m.get("anchors"):
            self.hr_anchor = True
        hr = self.hr
        self.pos0, self.sigma0, self.pos_end, self.sigma_end = _T(hr.pos0), _T(hr.sigma0), _T(hr.pos_end), _T(hr.sigma_end)
        self.t_end, self.n_obs, self.n_events = _T(hr.t_end), _T(hr.n_obs), _T(hr.n_events)
        self.status, self.draws_used = _T(hr.status), _T(hr.draws_used)
        self.n, self.n_end, self.bound_end, self.bound0 = _T(hr.n), _T(hr.n_end), None, None
        self.launches = []

    def run_replay(self, d, off, resume=None, max_events=0, spec_from=-1):
        hr = self.hr
        dr = np.ascontiguousarray(d.numpy()); of = np.ascontiguousarray(off.numpy())
        anch = hr.anchor_mask is not None
        b, keep = make_batch(1, hr.n_max, hr.M, record=7, spec_from=spec_from, max_events=max_events, exit_cap=hr.n_max if anch else 0,
                             anchor_mask=hr.anchor_mask, bound0=hr.bound_end if anch else None, bound_end=hr.bound_end if anch else None,
                             obs_bound=hr.obs_bound if anch else None, exit_t=hr.exit_t if anch else None,
                             exit_pos=hr.exit_pos if anch else None, n_exit=hr.n_exit if anch else None, n_end=hr.n_end, obs_n=hr.obs_n,
                             times_obs=hr.times_obs, weights=hr.weights, beta=hr.beta, n=self.n.a, pos0=self.pos0.a,
                             sigma0=self.sigma0.a, draws=dr, draw_off=of, t_start=hr.t_end, obs_start=hr.n_obs,
                             ev_start=hr.n_events, obs_cp=hr.obs_cp, obs_cm=hr.obs_cm, obs_pos=hr.obs_pos,
                             obs_sigma_sum=hr.obs_sigma_sum, obs_m_local=hr.obs_m_local, n_obs=hr.n_obs,
                             n_events=hr.n_events, t_end=hr.t_end, status=hr.status, draws_used=hr.draws_used,
                             pos_end=self.pos_end.a, sigma_end=self.sigma_end.a)
        assert oracle.load().aps_oracle_run(self.P, b, 0, 1) == 0
        self.launches.append((dr.size, int(hr.status[0]), int(hr.draws_used[0])))


def build(c, rng):
    import test_dropin_gpu as T
    return T.build(c, rng)


@pytest.mark.parametrize("name", ["c2_sweep_b0", "tiny_diffusive", "k1_dense", "crowding", "c1_exclusion", "anchors_k3",
                                  "anchors_crowding_global"])
def test_chunked_replay_with_rewind(name, monkeypatch):
    c = load_case(name)
    m = c["meta"]
    g = np.random.default_rng(m["seed"])
    ps = build(c, g)
    pos, sigma = ps.init_particles()
    rb = OracleBackedBatch(c, pos, sigma)
    ps._run_replay(rb)
    hr = rb.hr
    hr.draws_used[0] = len(c["draws"])     # per-launch counter; the log-level total is checked via the rng below
    hr.trace = None
    assert_matches_reference(c, hr)
    # Generator state == the reference's consumption (init + the recorded variates)
    h = np.random.default_rng(m["seed"])
    build(c, h).init_particles()
    for ev in range(m["n_events"]):
        h.exponential(1.0); h.random(); h.random()
        if c["trace"][ev, 1] < 2:
            h.random()
    assert g.random() == h.random()
    if name == "tiny_diffusive":
        assert len(rb.launches) > 5      # many diffusive events -> many rewinds


def test_event_by_event_path_for_duck_rng():
    c = load_case("tiny_diffusive")
    m = c["meta"]

    class Duck:
        def __init__(self, seed): self.g = np.random.default_rng(seed)
        def exponential(self, s=1.0): return self.g.exponential(s)
        def random(self): return self.g.random()
        def choice(self, *a, **k): return self.g.choice(*a, **k)
        def poisson(self, *a, **k): return self.g.poisson(*a, **k)

    ps = build(c, Duck(m["seed"]))
    pos, sigma = ps.init_particles()
    rb = OracleBackedBatch(c, pos, sigma)
    ps._run_replay(rb)
    rb.hr.draws_used[0] = len(c["draws"]); rb.hr.trace = None
    assert_matches_reference(c, rb.hr)
