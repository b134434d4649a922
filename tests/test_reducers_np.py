"""The numpy restatement of the drivers' reducers against the reference's own functions
(tests/golden/reducers.json was produced by executing them; tools/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

from common import GOLDEN, load_case
from oracle import reducers_np as rn


@pytest.mark.parametrize("tag", ["b0", "b2"])
def test_reducers_match_reference(tag):
    want = json.load(open(os.path.join(GOLDEN, "reducers.json")))[tag]
    c = load_case(f"reducers_{tag}")
    m = c["meta"]
    pos_list = [c["pos_obs"][k].astype(np.int64) for k in range(len(c["times_obs"]))]
    got = rn.all_reducers(c["times_obs"], c["rho_p_list"], c["rho_m_list"], c["total_list"], c["m_global"],
                          pos_list, m["L"], m["dx"])
    assert (got["si"], got["ei"]) == (want["si"], want["ei"])
    for k in ["mean_v", "D_eff", "m_mean", "rho_eff", "block"]:
        assert got[k] == pytest.approx(want[k], rel=1e-12, abs=1e-15), k
    np.testing.assert_allclose(got["v_eff"], want["v_eff"], rtol=1e-12, atol=1e-15)


def test_window_quirk_branches():
    """`~safe[start_idx:]` on an index array: non-empty slice -> window collapses to the minimum length."""
    L, M = 50, 40
    times = np.arange(M) * 0.1
    total = np.zeros((M, L)); total[:, 10] = 1.0
    assert rn.v_eff_and_window(times, total, L)[2:4] == (26, 40)          # nothing near the boundary
    total[:, L - 1] = 5.0                                                  # every row "unsafe" (40 > 26)
    assert rn.v_eff_and_window(times, total, L)[2:4] == (26, 30)
    total[:, L - 1] = 0.0; total[:5, L - 1] = 5.0                          # 5 unsafe rows <= start_idx
    assert rn.v_eff_and_window(times, total, L)[2:4] == (26, 40)


def test_gaussian_weights_equal_scipy():
    from scipy.ndimage import _filters
    from aps_b200.engine import gaussian_weights
    for sd in [0.05, 0.5, 1.6, 2.0, 5.0, 20.0, 300.0]:
        r, w = gaussian_weights(sd)
        assert r == int(4.0 * sd + 0.5)
        assert np.array_equal(w, _filters._gaussian_kernel1d(sd, 0, r)[::-1])
