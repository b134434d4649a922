"""ctypes binding of the C ABI declared in include/aps.h.

The structures here are the Python spelling of `aps_params` / `aps_batch`; the oracle's test
binding (oracle/oracle.py) re-uses them so that both sides are driven with the same descriptor.
The product library is `csrc/libaps_b200.so`, built in-tree by `build.py`.  There is no CPU
fallback: if the library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "csrc", "libaps_b200.so")

ABI_VERSION = 5          # APS_ABI_VERSION of include/aps.h
APS_OK = 0
APS_ERR_INVALID, APS_ERR_NO_DEVICE, APS_ERR_CUDA, APS_ERR_CAPACITY = 1, 2, 3, 4
APS_RUN_DONE, APS_RUN_EMPTY, APS_RUN_DRAWS_EXHAUSTED, APS_RUN_MAX_EVENTS = 0, 1, 2, 3
APS_FLAG_CROWDING, APS_FLAG_SUPPRESS_FLIP_BOUND, APS_FLAG_IMMOBILIZE, APS_FLAG_PERIODIC = 1, 2, 4, 8
APS_REC_COUNTS, APS_REC_POS, APS_REC_MLOCAL = 1, 2, 4
APS_EV_DIFF_LEFT, APS_EV_DIFF_RIGHT, APS_EV_ACTIVE, APS_EV_FLIP = 0, 1, 2, 3
APS_EV_BIND, APS_EV_UNBIND, APS_EV_EXIT = 4, 5, 6


class ApsParams(C.Structure):
    _fields_ = [
        ("L", C.c_int32),
        ("K", C.c_int32),
        ("radius", C.c_int32),
        ("flags", C.c_uint32),
        ("rate_diffusion", C.c_double),
        ("rate_active", C.c_double),
        ("T", C.c_double),
        ("k_on", C.c_double),
        ("k_off", C.c_double),
        ("k_exit", C.c_double),
    ]


class ApsBatch(C.Structure):
    _fields_ = [
        ("n_replicas", C.c_int32),
        ("n_max", C.c_int32),
        ("M", C.c_int32),
        ("record", C.c_uint32),
        ("max_events", C.c_int64),
        ("trace_cap", C.c_int64),
        ("spec_from", C.c_int64),
        ("times_obs", C.c_void_p),
        ("weights", C.c_void_p),
        ("beta", C.c_void_p),
        ("n", C.c_void_p),
        ("pos0", C.c_void_p),
        ("sigma0", C.c_void_p),
        ("draws", C.c_void_p),
        ("draw_off", C.c_void_p),
        ("seeds", C.c_void_p),
        ("t_start", C.c_void_p),
        ("obs_start", C.c_void_p),
        ("ev_start", C.c_void_p),
        ("obs_cp", C.c_void_p),
        ("obs_cm", C.c_void_p),
        ("obs_pos", C.c_void_p),
        ("obs_sigma_sum", C.c_void_p),
        ("obs_m_local", C.c_void_p),
        ("n_obs", C.c_void_p),
        ("n_events", C.c_void_p),
        ("t_end", C.c_void_p),
        ("status", C.c_void_p),
        ("n_guard", C.c_void_p),
        ("draws_used", C.c_void_p),
        ("pos_end", C.c_void_p),
        ("sigma_end", C.c_void_p),
        ("trace", C.c_void_p),
        ("m_field_in", C.c_void_p),
        ("anchor_mask", C.c_void_p),
        ("bound0", C.c_void_p),
        ("n_end", C.c_void_p),
        ("bound_end", C.c_void_p),
        ("obs_n", C.c_void_p),
        ("obs_bound", C.c_void_p),
        ("exit_t", C.c_void_p),
        ("exit_pos", C.c_void_p),
        ("n_exit", C.c_void_p),
        ("exit_cap", C.c_int64),
        ("flip_tab", C.c_void_p),
        ("flip_G", C.c_int64),
        ("weights_host", C.c_void_p),
    ]


class ApsInitArgs(C.Structure):
    _fields_ = [("n_replicas", C.c_int32), ("L", C.c_int32), ("K", C.c_int32), ("n_max", C.c_int32),
                ("mode", C.c_int32), ("N_fixed", C.c_int32), ("n_profiles", C.c_int32), ("reserved", C.c_int32),
                ("rho0_plus", C.c_void_p), ("rho0_minus", C.c_void_p), ("profile_of", C.c_void_p),
                ("N_of", C.c_void_p), ("seeds", C.c_void_p), ("pos0", C.c_void_p), ("sigma0", C.c_void_p),
                ("n", C.c_void_p)]


class ApsExpandArgs(C.Structure):
    _fields_ = [("n_replicas", C.c_int32), ("M", C.c_int32), ("L", C.c_int32), ("reserved", C.c_int32),
                ("dx", C.c_double), ("n", C.c_void_p), ("n_obs", C.c_void_p), ("obs_cp", C.c_void_p),
                ("obs_cm", C.c_void_p), ("rho_p", C.c_void_p), ("rho_m", C.c_void_p), ("total", C.c_void_p),
                ("var", C.c_void_p), ("obs_n", C.c_void_p)]


APS_RED_V_EFF, APS_RED_D_EFF, APS_RED_M_MEAN, APS_RED_RHO_EFF, APS_RED_BLOCK = 0, 1, 2, 3, 4
APS_RED_START, APS_RED_END, APS_RED_NOBS, APS_RED_N = 5, 6, 7, 8


class ApsReduceArgs(C.Structure):
    _fields_ = [("n_replicas", C.c_int32), ("M", C.c_int32), ("L", C.c_int32), ("n_max", C.c_int32),
                ("dx", C.c_double), ("boundary_xmin", C.c_double), ("max_boundary_fraction", C.c_double),
                ("min_window_fraction", C.c_double), ("window_fraction", C.c_double),
                ("times_obs", C.c_void_p), ("n", C.c_void_p), ("n_obs", C.c_void_p), ("obs_cp", C.c_void_p),
                ("obs_cm", C.c_void_p), ("obs_pos", C.c_void_p), ("obs_sigma_sum", C.c_void_p),
                ("out", C.c_void_p), ("v_eff", C.c_void_p)]


class ApsProfileArgs(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("reps_per_point", C.c_int32), ("M", C.c_int32), ("L", C.c_int32),
                ("row_lo", C.c_int32), ("row_hi", C.c_int32), ("dx", C.c_double), ("n", C.c_void_p),
                ("n_obs", C.c_void_p), ("obs_cp", C.c_void_p), ("obs_cm", C.c_void_p), ("prof", C.c_void_p),
                ("point_start", C.c_void_p), ("point_reps", C.c_void_p), ("scratch", C.c_void_p),
                ("n_replicas", C.c_int32), ("reserved", C.c_int32)]


class ApsHistArgs(C.Structure):
    _fields_ = [("n_replicas", C.c_int32), ("M", C.c_int32), ("n_points", C.c_int32), ("n_bins", C.c_int32),
                ("row_lo", C.c_int32), ("row_hi", C.c_int32), ("accumulate", C.c_int32), ("reserved", C.c_int32),
                ("lo", C.c_double), ("hi", C.c_double),
                ("n", C.c_void_p), ("n_obs", C.c_void_p), ("obs_sigma_sum", C.c_void_p), ("obs_n", C.c_void_p),
                ("point_of", C.c_void_p), ("mbar", C.c_void_p), ("hist", C.c_void_p)]


APS_K2_MAX_TRIALS = 64


class ApsK2Rates(C.Structure):
    _fields_ = [("t_left", C.c_uint32), ("t_right", C.c_uint32), ("t_active", C.c_uint32), ("n_cdf", C.c_uint32),
                ("inv_cmax", C.c_double), ("beta", C.c_double), ("mu", C.c_double),
                ("cdf32", C.c_uint32 * APS_K2_MAX_TRIALS)]


class ApsK2Args(C.Structure):
    _fields_ = [("L", C.c_int64), ("L_global", C.c_int64), ("global_offset", C.c_int64), ("n_particles", C.c_int64),
                ("seed", C.c_uint64), ("pass_", C.c_uint64), ("radius", C.c_int32), ("reserved", C.c_int32),
                ("rates", ApsK2Rates), ("w16", C.c_void_p), ("flip_tab", C.c_void_p), ("in_", C.c_void_p), ("out", C.c_void_p),
                ("msum_in", C.c_void_p), ("msum_out", C.c_void_p), ("count_lo", C.c_int64), ("count_hi", C.c_int64)]


class ApsPdeArgs(C.Structure):          # include/aps_pde.h
    _fields_ = [("L", C.c_int32), ("n_runs", C.c_int32), ("bc", C.c_int32), ("model", C.c_int32), ("field", C.c_int32),
                ("snapshot_interval", C.c_int32), ("n_tracers", C.c_int32), ("window", C.c_int32), ("nsteps", C.c_int64),
                ("dt", C.c_double), ("dx", C.c_double), ("xlim", C.c_double)] + \
               [(k, C.c_void_p) for k in ["beta", "lam", "gamma", "kernel", "radius", "seeds", "rho_p", "rho_m", "m_series",
                                          "var_series", "snapshots", "m_snapshots", "tracer_pos", "tracer_state",
                                          "tracer_hist", "v_eff_series", "D_eff_series", "tot_series"]]


class ApsK2Multi(C.Structure):         # include/aps.h aps_k2_multi
    _fields_ = [("n_passes", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("refresh_every", C.c_int32),
                ("ghost", C.c_int64), ("own_lo", C.c_int64), ("own_hi", C.c_int64), ("buf0", C.c_void_p), ("buf1", C.c_void_p),
                ("sync", C.c_void_p), ("peer", C.c_void_p * 8)]


APS_PDE_BC = {"periodic": 0, "neumann": 1}
APS_PDE_MODEL = {"bidirectional": 0, "anchored_minus": 1}

# every symbol include/aps.h and include/aps_pde.h declare: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "aps_abi_version": (C.c_int, []),
    "aps_last_error": (C.c_char_p, []),
    "aps_device_count": (C.c_int, []),
    "aps_set_device": (C.c_int, [C.c_int]),
    "aps_run_replay_device": (C.c_int, [_P(ApsParams), _P(ApsBatch), C.c_void_p]),
    "aps_run_philox_device": (C.c_int, [_P(ApsParams), _P(ApsBatch), C.c_void_p]),
    "aps_run_replay_host": (C.c_int, [_P(ApsParams), _P(ApsBatch)]),
    "aps_run_philox_host": (C.c_int, [_P(ApsParams), _P(ApsBatch)]),
    "aps_launch_count": (C.c_int64, []),
    "aps_replica_smem_bytes": (C.c_int64, [_P(ApsParams), C.c_int32]),
    "aps_init_particles_device": (C.c_int, [_P(ApsInitArgs), C.c_void_p]),
    "aps_m_field_host": (C.c_int, [_P(ApsParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aps_expand_obs_device": (C.c_int, [_P(ApsExpandArgs), C.c_void_p]),
    "aps_reduce_runs_device": (C.c_int, [_P(ApsReduceArgs), C.c_void_p]),
    "aps_profile_sums_device": (C.c_int, [_P(ApsProfileArgs), C.c_void_p]),
    "aps_m_histogram_device": (C.c_int, [_P(ApsHistArgs), C.c_void_p]),
    "aps_k2_rates_init": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_double, _P(ApsK2Rates)]),
    "aps_k2_flip_table": (C.c_int, [_P(ApsK2Rates), C.c_void_p]),
    "aps_k2_pass_device": (C.c_int, [_P(ApsK2Args), C.c_void_p]),
    "aps_k2_run_device": (C.c_int, [_P(ApsK2Args), C.c_int, C.c_void_p]),
    "aps_k2_run_persistent_device": (C.c_int, [_P(ApsK2Args), _P(ApsK2Multi), C.c_void_p]),
    "aps_k2_peer_region_bytes": (C.c_int, []),
    "aps_k2_peer_alloc": (C.c_int, [_P(C.c_void_p), C.c_void_p]),
    "aps_k2_peer_open": (C.c_int, [C.c_void_p, _P(C.c_void_p)]),
    "aps_k2_peer_close": (C.c_int, [C.c_void_p]),
    "aps_k2_peer_free": (C.c_int, [C.c_void_p]),
    "aps_k2_init_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_void_p]),
    "aps_k2_profile_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aps_pde_solve_device": (C.c_int, [_P(ApsPdeArgs), C.c_void_p]),
    "aps_pde_smem_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "aps_debug_set_guard_scale": (None, [C.c_double]),
    "aps_debug_set_k1_threads": (None, [C.c_int]),
    "aps_debug_set_use_lut": (None, [C.c_int]),
    "aps_debug_set_use_fast": (None, [C.c_int]),
    "aps_debug_set_k2_ctas_per_sm": (None, [C.c_int]),
    "aps_debug_set_k2_stash_cap": (None, [C.c_int]),
    "aps_debug_set_reduce_impl": (None, [C.c_int]),
    "aps_debug_set_reduce_threads": (None, [C.c_int]),
}

_lib = None


class ApsError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero aps_status."""


def load(path: str | None = None):
    """Load libaps_b200.so and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ApsError(
            f"{p} is missing: build it with `python __graft_entry__.py build` "
            "(this package has no CPU fallback)"
        )
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.aps_abi_version() != ABI_VERSION:
        raise ApsError("libaps_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = "aps call"):
    if rc != APS_OK:
        msg = load().aps_last_error()
        raise ApsError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
