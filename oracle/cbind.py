"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE, NOT PRODUCT CODE)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gwas_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_colstats.argtypes = [dp, i64, i64, i64, dp, dp]
        L.oracle_gwasols_raw.argtypes = [dp, i64, i64, i64, dp, dp, dp, dp, ctypes.POINTER(ctypes.c_uint8)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(n)
    return n


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def colstats(A: np.ndarray):
    A = np.asfortranarray(A, dtype=np.float64)
    n, p = A.shape
    mean = np.empty(p)
    sd = np.empty(p)
    lib().oracle_colstats(_dp(A), n, p, n, _dp(mean), _dp(sd))
    return mean, sd


def gwasols_raw(A: np.ndarray, ys: np.ndarray, pc: np.ndarray):
    """(stat_ols, stat_lmm, keep) following gwas.jl:112-113, :129, :241-245 per marker."""
    A = np.asfortranarray(A, dtype=np.float64)
    n, p = A.shape
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    pc = np.ascontiguousarray(pc, dtype=np.float64)
    so = np.empty(p)
    sl = np.empty(p)
    keep = np.empty(p, dtype=np.uint8)
    lib().oracle_gwasols_raw(_dp(A), n, p, n, _dp(ys), _dp(pc), _dp(so), _dp(sl),
                             keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return so, sl, keep.astype(bool)
