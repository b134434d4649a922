"""The CPU oracle against the golden vectors produced by the unmodified reference
(tools/gen_golden.py): event-by-event trace, every observation row, bit-for-bit."""
import numpy as np
import pytest

from common import (assert_matches_reference, case_names, hostrun_from_case, load_case, params_from_case,
                    run_oracle)


@pytest.mark.parametrize("name", case_names())
def test_oracle_replay_matches_reference(name):
    c = load_case(name)
    hr = run_oracle(params_from_case(c), hostrun_from_case(c))
    assert_matches_reference(c, hr)


def test_oracle_threads_do_not_change_results():
    c = load_case("k1_dense")
    from common import HostRun
    m = c["meta"]
    R = 5
    n = m["n"]
    dr = np.tile(c["draws"], R)
    off = np.arange(R + 1) * len(c["draws"])
    def mk():
        return HostRun(m["L"], n, len(c["times_obs"]), [n] * R, np.tile(c["pos0"], R), np.tile(c["sigma0"], R),
                       [m["ps"]["beta"]] * R, c["times_obs"], c["weights"], draws=dr, draw_off=off)
    a = run_oracle(params_from_case(c), mk(), threads=1)
    b = run_oracle(params_from_case(c), mk(), threads=3)
    for f in ["obs_cp", "obs_pos", "n_events", "t_end"]:
        assert np.array_equal(getattr(a, f), getattr(b, f))
    for r in range(R):
        assert_matches_reference(c, a, rep=r)
