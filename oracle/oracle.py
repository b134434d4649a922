"""ctypes binding of oracle/libaps_oracle.so (TEST INFRASTRUCTURE ONLY — see aps_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from aps_b200.capi import ApsBatch, ApsInitArgs, ApsK2Args, ApsK2Rates, ApsParams  # noqa: E402  (shared descriptor layout)

LIB = os.path.join(HERE, "libaps_oracle.so")
_lib = None


def build(force: bool = False):
    src = os.path.join(HERE, "aps_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B" if force else "-s"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.aps_oracle_run.restype = C.c_int
        lib.aps_oracle_run.argtypes = [C.POINTER(ApsParams), C.POINTER(ApsBatch), C.c_int, C.c_int]
        lib.aps_oracle_pairwise_sum.restype = C.c_double
        lib.aps_oracle_pairwise_sum.argtypes = [C.c_void_p, C.c_int64]
        lib.aps_oracle_exp.restype = C.c_double
        lib.aps_oracle_exp.argtypes = [C.c_double]
        lib.aps_oracle_log.restype = C.c_double
        lib.aps_oracle_log.argtypes = [C.c_double]
        lib.aps_oracle_philox.restype = None
        lib.aps_oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.aps_oracle_m_field.restype = C.c_int
        lib.aps_oracle_m_field.argtypes = [C.POINTER(ApsParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.aps_oracle_init.restype = C.c_int
        lib.aps_oracle_init.argtypes = [C.POINTER(ApsInitArgs)]
        lib.aps_oracle_k2_pass.restype = C.c_int
        lib.aps_oracle_k2_pass.argtypes = [C.POINTER(ApsK2Args)]
        lib.aps_oracle_k2_run.restype = C.c_int
        lib.aps_oracle_k2_run.argtypes = [C.POINTER(ApsK2Args), C.c_int]
        lib.aps_oracle_k2_rates.restype = C.c_int
        lib.aps_oracle_k2_rates.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(ApsK2Rates)]
        lib.aps_oracle_philox_log.restype = C.c_int64
        lib.aps_oracle_philox_log.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        lib.aps_oracle_k2_flip_table.restype = None
        lib.aps_oracle_k2_flip_table.argtypes = [C.c_void_p, C.c_void_p]
        lib.aps_oracle_k2_init.restype = None
        lib.aps_oracle_k2_init.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_double, C.c_double]
        _lib = lib
    return _lib
