"""Operand loading -- host-side mirror of the reference's src/py2rust.rs entry points.

The reference runs two small embedded Python snippets under the GIL (py2rust.rs:6-44 for
``.pkl`` dictionaries, :63-81 for Matrix Market files).  The functions below keep their names,
arguments, error behaviour and stdout lines, and use the same scipy conversions, so the
operands -- and the CLI transcript -- match the reference:

* ``.mtx``: ``scipy.io.mmread(file).tocsr()``; array-format files make ``mmread`` return an
  ndarray, whose missing ``.tocsr`` raises AttributeError (the reference panics on it).
* ``.pkl``: ``pickle.load(f)[name] -> (A, B)``; csc/coo -> ``tocsr()``, ndarray ->
  ``csr_matrix``, csr passes through, anything else -> TypeError.

One documented deviation: a pickled csr_matrix that is not canonical is
canonicalised (duplicates summed, columns sorted) instead of passed through unchecked; the
engine -- like the reference's sort/merge units -- needs ascending unique columns.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import scipy.io
import scipy.sparse as sp

from .gemm import GEMM, _canonical

_BANNER = "---- Python Interface ----"


def _as_csr(x) -> sp.csr_matrix:
    if isinstance(x, sp.csr_matrix):
        return x
    if isinstance(x, (sp.csc_matrix, sp.coo_matrix)):
        return x.tocsr()
    if isinstance(x, np.ndarray):
        return sp.csr_matrix(x)
    raise TypeError("Unsupported matrix type: {}".format(type(x)))


def _describe(tag: str, m: sp.csr_matrix) -> None:
    print(f"% -- {tag} --")
    print(f"% shape: {m.shape} data: {m.data[:5]}... indices: {m.indices[:5]}... indptr: {m.indptr[:5]}...")


def load_pickled_gemms(gemm_fp: str, gemm_nm: str) -> GEMM:
    """NN workloads: the (A, B) pair stored under ``gemm_nm`` in the pickle ``gemm_fp``."""
    print(_BANNER)
    print(f"% Load {gemm_nm} from", gemm_fp)
    with open(gemm_fp, "rb") as fh:
        pair = pickle.load(fh)[gemm_nm]
    a, b = (_canonical(_as_csr(x)) for x in pair)
    _describe("A", a)
    _describe("B", b)
    print("--- Return from Python Interface ---\n")
    return GEMM.new(gemm_nm, (a.shape, a.indptr, a.indices, a.data, b.shape, b.indptr, b.indices, b.data))


def load_mm_mat(dir_path: str, gemm_nm: str) -> sp.csr_matrix:
    """SS workloads: ``<dir_path>/<gemm_nm>.mtx`` as canonical CSR with f64 values."""
    print(_BANNER)
    print(f"% Load {gemm_nm} from {dir_path}")
    with open(os.path.join(dir_path, gemm_nm + ".mtx"), "r") as fh:
        mat = scipy.io.mmread(fh).tocsr()
    return _canonical(mat)
