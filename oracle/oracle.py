"""ctypes front-end of oracle/liboracle.so (CPU restatement, see spgemm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package (spada-sim_b200/), which has no CPU fallback.

Parity status: "parity unpinned" by the reference (it has no tests); pinned here
against scipy and the survey's cari known answers (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "spgemm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        i64, i32p, i64p, f64p = C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        _LIB.oracle_max_threads.restype = C.c_int
        _LIB.oracle_flops.restype = i64
        _LIB.oracle_flops.argtypes = [i64, i64p, i32p, i64p, i64p]
        _LIB.oracle_spgemm_symbolic.restype = i64
        _LIB.oracle_spgemm_symbolic.argtypes = [i64, i64, i64p, i32p, i64p, i32p, i64p, C.c_int]
        _LIB.oracle_spgemm_numeric.restype = None
        _LIB.oracle_spgemm_numeric.argtypes = [i64, i64, i64p, i32p, f64p, i64p, i32p, f64p, i64p, i32p, f64p, C.c_int]
        _LIB.oracle_transpose.restype = None
        _LIB.oracle_transpose.argtypes = [i64, i64, i64p, i32p, f64p, i64p, i32p, f64p]
        _LIB.oracle_validate_csr.restype = i64
        _LIB.oracle_validate_csr.argtypes = [i64, i64, i64p, i32p]
        _LIB.oracle_parse_group.restype = i64
        _LIB.oracle_parse_group.argtypes = [i64, i64p, C.c_float, i64p]
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _arrays(m: sp.csr_matrix):
    return (np.ascontiguousarray(m.indptr, dtype=np.int64),
            np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=np.float64))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def flops(a: sp.csr_matrix, b: sp.csr_matrix) -> np.ndarray:
    ap, aj, _ = _arrays(a)
    bp, _, _ = _arrays(b)
    out = np.zeros(a.shape[0], dtype=np.int64)
    lib().oracle_flops(a.shape[0], _p(ap, C.c_int64), _p(aj, C.c_int32), _p(bp, C.c_int64), _p(out, C.c_int64))
    return out


def spgemm(a: sp.csr_matrix, b: sp.csr_matrix, threads: int = 1):
    """C = A x B, canonical CSR.  Returns (indptr int64, indices int32, data float64)."""
    if a.shape[1] != b.shape[0]:
        raise ValueError("dimension mismatch")
    m, n = a.shape[0], b.shape[1]
    ap, aj, ax = _arrays(a)
    bp, bj, bx = _arrays(b)
    cp = np.zeros(m + 1, dtype=np.int64)
    nnz = lib().oracle_spgemm_symbolic(m, n, _p(ap, C.c_int64), _p(aj, C.c_int32), _p(bp, C.c_int64),
                                       _p(bj, C.c_int32), _p(cp, C.c_int64), threads)
    cj = np.empty(nnz, dtype=np.int32)
    cx = np.empty(nnz, dtype=np.float64)
    lib().oracle_spgemm_numeric(m, n, _p(ap, C.c_int64), _p(aj, C.c_int32), _p(ax, C.c_double),
                                _p(bp, C.c_int64), _p(bj, C.c_int32), _p(bx, C.c_double),
                                _p(cp, C.c_int64), _p(cj, C.c_int32), _p(cx, C.c_double), threads)
    return cp, cj, cx


def spgemm_csr(a, b, threads: int = 1) -> sp.csr_matrix:
    cp, cj, cx = spgemm(a, b, threads)
    return sp.csr_matrix((cx, cj, cp), shape=(a.shape[0], b.shape[1]))


def transpose(a: sp.csr_matrix) -> sp.csr_matrix:
    m, n = a.shape
    ap, aj, ax = _arrays(a)
    bp = np.zeros(n + 1, dtype=np.int64)
    bj = np.empty(a.nnz, dtype=np.int32)
    bx = np.empty(a.nnz, dtype=np.float64)
    lib().oracle_transpose(m, n, _p(ap, C.c_int64), _p(aj, C.c_int32), _p(ax, C.c_double),
                           _p(bp, C.c_int64), _p(bj, C.c_int32), _p(bx, C.c_double))
    return sp.csr_matrix((bx, bj, bp), shape=(n, m))


def validate_csr(a: sp.csr_matrix) -> int:
    ap, aj, _ = _arrays(a)
    return int(lib().oracle_validate_csr(a.shape[0], a.shape[1], _p(ap, C.c_int64), _p(aj, C.c_int32)))


def parse_group(a: sp.csr_matrix, var_factor: float = 1.5) -> np.ndarray:
    ap, _, _ = _arrays(a)
    out = np.zeros(a.shape[0] + 1, dtype=np.int64)
    n = lib().oracle_parse_group(a.shape[0], _p(ap, C.c_int64), var_factor, _p(out, C.c_int64))
    return out[:n].copy()
