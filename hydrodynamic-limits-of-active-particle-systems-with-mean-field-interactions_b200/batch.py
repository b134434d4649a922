"""Assemble `aps_batch` descriptors from numpy arrays (host pointers) or torch tensors
(device pointers).  Pure plumbing: checks dtypes/contiguity/shapes and keeps the buffers alive."""
from __future__ import annotations

from .capi import ApsBatch, ApsParams

# field -> (dtype name, required?)
_FIELDS = {
    "times_obs": "float64",
    "weights": "float64",
    "beta": "float64",
    "n": "int32",
    "pos0": "int32",
    "sigma0": "int8",
    "draws": "float64",
    "draw_off": "int64",
    "seeds": "uint64",
    "t_start": "float64",
    "obs_start": "int32",
    "ev_start": "int64",
    "obs_cp": "int8",
    "obs_cm": "int8",
    "obs_pos": "int32",
    "obs_sigma_sum": "int32",
    "obs_m_local": "float64",
    "n_obs": "int32",
    "n_events": "int64",
    "t_end": "float64",
    "status": "int32",
    "n_guard": "int64",
    "draws_used": "int64",
    "pos_end": "int32",
    "sigma_end": "int8",
    "trace": "int32",
    "m_field_in": "float64",
    "anchor_mask": "uint8",
    "bound0": "int8",
    "n_end": "int32",
    "bound_end": "int8",
    "obs_n": "int32",
    "obs_bound": "int8",
    "exit_t": "float64",
    "exit_pos": "int32",
    "n_exit": "int32",
    "flip_tab": "float64",
    "weights_host": "float64",
}


def _dtype_name(a) -> str:
    return str(a.dtype).replace("torch.", "")


def _ptr(a):
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def _contig(a) -> bool:
    if hasattr(a, "is_contiguous"):
        return a.is_contiguous()
    return a.flags["C_CONTIGUOUS"]


def make_params(L, K, radius, D, lam, T, flags=0, k_on=0.0, k_off=0.0, k_exit=0.0) -> ApsParams:
    return ApsParams(int(L), int(K), int(radius), int(flags), float(D), float(lam), float(T), float(k_on), float(k_off),
                     float(k_exit))


def make_batch(n_replicas, n_max, M, record=0, max_events=0, trace_cap=0, spec_from=-1, exit_cap=0, flip_G=0, **arrays):
    """Returns (ApsBatch, keepalive list).  uint64 seeds may be passed as int64 torch tensors
    (torch has limited uint64 support); the bit pattern is what matters."""
    b = ApsBatch()
    b.n_replicas, b.n_max, b.M = int(n_replicas), int(n_max), int(M)
    b.record, b.max_events, b.trace_cap = int(record), int(max_events), int(trace_cap)
    b.spec_from = int(spec_from)
    b.exit_cap = int(exit_cap)
    b.flip_G = int(flip_G)
    keep = []
    for name, arr in arrays.items():
        if name not in _FIELDS:
            raise KeyError(f"unknown aps_batch field {name!r}")
        if arr is None:
            continue
        want = _FIELDS[name]
        have = _dtype_name(arr)
        if have != want and not (want == "uint64" and have == "int64"):
            raise TypeError(f"aps_batch.{name}: expected {want}, got {have}")
        if not _contig(arr):
            raise ValueError(f"aps_batch.{name} must be contiguous")
        setattr(b, name, _ptr(arr))
        keep.append(arr)
    return b, keep
