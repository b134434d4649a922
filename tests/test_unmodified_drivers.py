"""The reference's driver scripts, byte for byte as shipped, executed end to end through the drop-in:
`from PARTICLE_solver_CLASS import ParticleSystem` resolves to dropin/PARTICLE_solver_CLASS.py (first on sys.path),
every `ParticleSystem(...).run()` of the script runs the K1 / K4 CUDA kernels through the C ABI, and the scripts'
own post-processing, fits, npz writers and plotting calls (`ps.plot_individuals`, PARTICLE_solver_BIOLOGY_EXCLUSION.py:107)
consume the returned dict unchanged.

The scripts are taken from baseline/_ref (tools/install_reference.py: an unmodified install of the reference that
travels to the GPU box) or /root/reference; the test is skipped where neither exists.  matplotlib / vispy are not
installed in this image: tests/plot_stubs.py stands in for them (the scripts import them at module top).
"""
import os
import runpy
import sys

import numpy as np
import pytest

from common import GOLDEN

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(GOLDEN))
DROPIN = os.path.join(ROOT, "dropin")


def _ref_dir():
    for d in (os.environ.get("APS_REFERENCE_PATH"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.exists(os.path.join(d, "PARTICLE_solver_BIOLOGY_EXCLUSION.py")):
            return d
    pytest.skip("no copy of the reference's driver scripts on this box (baseline/_ref or /root/reference)")


def _run_script(name, tmp_path, monkeypatch):
    import plot_stubs
    plot_stubs.install()
    ref = _ref_dir()
    monkeypatch.chdir(tmp_path)                                   # the scripts write their png / npz files into the cwd
    monkeypatch.syspath_prepend(DROPIN)
    for m in ("PARTICLE_solver_CLASS", "IMEX_PDE_solver_class"):
        mod = sys.modules.get(m)
        if mod is not None and not os.path.abspath(mod.__file__).startswith(DROPIN):
            monkeypatch.delitem(sys.modules, m)
    from aps_b200 import capi
    n0 = capi.load().aps_launch_count()
    g = runpy.run_path(os.path.join(ref, name), run_name="__main__")
    import PARTICLE_solver_CLASS as C
    assert os.path.abspath(C.__file__).startswith(DROPIN), "the script imported something else than the drop-in"
    assert g["ParticleSystem"] is C.ParticleSystem
    assert capi.load().aps_launch_count() > n0, "no CUDA kernel was launched"
    return g


def test_single_run_driver_with_plots(tmp_path, monkeypatch):
    """PARTICLE_solver_BIOLOGY_EXCLUSION.py (BASELINE config 1): constructor :55-94, run :95-97, plot_individuals :107."""
    g = _run_script("PARTICLE_solver_BIOLOGY_EXCLUSION.py", tmp_path, monkeypatch)
    out, ps = g["out"], g["ps"]
    assert out["rho_p_list"].shape == (40, 1000) and out["fft_amp_list"].shape == (40, 1000)
    assert all(p is not None and p.size == 750 for p in out["pos_list"])
    assert ps.last_run_info["mode"] == "philox" and ps.last_run_info["n_events"] > 30_000
    assert (out["total_list"].sum(axis=1) * ps.dx).round(9).tolist() == [1.0] * 40        # CLASS.py:209-213 normalisation


def test_beta_sweep_driver_writes_the_reference_npz(tmp_path, monkeypatch):
    """PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py: 11 betas x 3 runs at import (:1030-1034), reducers, fits, npz."""
    g = _run_script("PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py", tmp_path, monkeypatch)
    keys = {"beta_values", "means", "stds", "ses", "D_means", "D_ses", "m_means", "m_stds", "m_ses", "rho_means", "rho_ses",
            "block_means", "block_ses", "ps_kwargs", "outs"}
    d = np.load(tmp_path / "CHANGES_simulation_out_sweep.npz", allow_pickle=True)
    assert set(d.files) == keys                                                          # sweep_beta.py:952-970
    assert os.path.exists(tmp_path / "CHANGES_after_simulation_out_sweep.npz")            # :1033-1034
    assert d["means"].shape == (11,) and np.isfinite(d["means"]).all() and d["outs"].shape[:2] == (11, 3)     # all 33 out dicts
    m = d["m_means"]
    assert abs(m[0]) < 0.25 and m[-1] > 0.45 and m[-1] > m[0] + 0.4      # m ~ 0 at beta = 0, ordering at beta = 3 (:232-254)
    assert g["save_dict"]["means"].shape == (11,)


def test_local_structure_driver(tmp_path, monkeypatch):
    """PARTICLE_solver_BIOLOGY_local_structure.py (BASELINE config 3 parameters, `__main__` block :671-753)."""
    g = _run_script("PARTICLE_solver_BIOLOGY_local_structure.py", tmp_path, monkeypatch)
    assert os.path.exists(tmp_path / "beta_sweep_local_structure.npz")
    res = g["results"]
    assert len(res) == 11 and all("var_mean" in v and "fft_mean_mean" in v for v in res.values())


def test_pde_run_driver_with_plots(tmp_path, monkeypatch):
    """IMEX_PDE_solver_run.py (the config-5 comparison run: L=1000, T=20, dt=5e-4, beta=2, sigma=0.005, seed=58, :7-27)
    through dropin/IMEX_PDE_solver_class.py: solve(), get_output(), plot_all(), plot_individual() (:29-34)."""
    import plot_stubs
    plot_stubs.install()
    ref = _ref_dir()
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(DROPIN)
    mod = sys.modules.get("IMEX_PDE_solver_class")
    if mod is not None and not os.path.abspath(mod.__file__).startswith(DROPIN):
        monkeypatch.delitem(sys.modules, "IMEX_PDE_solver_class")
    g = runpy.run_path(os.path.join(ref, "IMEX_PDE_solver_run.py"), run_name="__main__")
    import IMEX_PDE_solver_class as C
    assert os.path.abspath(C.__file__).startswith(DROPIN) and g["IMEXPDE"] is C.IMEXPDE
    out = g["out"]
    assert out["fft_amp"].shape == (40001, 501) and out["m_series"].shape == (40001,)
    assert out["snapshots"].shape == (801, 1000) and abs((out["rho_p"] + out["rho_m"]).sum() - 1.0) < 1e-9
    assert os.path.isdir(tmp_path / "IMEX_beta_3p0")
