"""GEMM operand pair -- mirror of the reference's src/gemm.rs.

``GEMM.from_mat`` (gemm.rs:41-53): square matrix => A x A, otherwise A x A^T with the
transpose materialised as CSR.  ``__str__`` reproduces the reference's Display impl
(gemm.rs:56-91) *including* its quirk of printing A's data/indices under "--B"
(gemm.rs:79, 84) so CLI output stays comparable line by line.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .rustfmt import debug_list


def _canonical(m) -> sp.csr_matrix:
    m = sp.csr_matrix(m)
    if not m.has_canonical_format:
        m = m.copy()
        m.sum_duplicates()
        m.sort_indices()
    if m.data.dtype != np.float64:
        m = m.astype(np.float64)  # pyo3 extracts every element as f64 (gemm.rs:23)
    return m


class GEMM:
    def __init__(self, name: str, a: sp.csr_matrix, b: sp.csr_matrix):
        self.name = name
        self.a = a
        self.b = b

    @classmethod
    def new(cls, name: str, raw) -> "GEMM":
        """gemm.rs:33-39: raw = (shape_A, indptr_A, indices_A, data_A, shape_B, indptr_B, indices_B, data_B)."""
        sa, pa, ia, da, sb, pb, ib, db = raw
        a = sp.csr_matrix((np.asarray(da, dtype=np.float64), np.asarray(ia), np.asarray(pa)), shape=tuple(sa))
        b = sp.csr_matrix((np.asarray(db, dtype=np.float64), np.asarray(ib), np.asarray(pb)), shape=tuple(sb))
        return cls(name, a, b)

    @classmethod
    def from_mat(cls, name: str, mat: sp.csr_matrix) -> "GEMM":
        """gemm.rs:41-53."""
        mat = _canonical(mat)
        if mat.shape[0] == mat.shape[1]:
            b = mat  # the reference clones; the engine uploads shared arrays once
        else:
            b = _canonical(mat.T)
        return cls(name, mat, b)

    def __str__(self) -> str:
        a, b = self.a, self.b
        na = min(len(a.data), 5)
        nb_d = min(len(b.data), 5)
        nb_i = min(len(b.indices), 5)
        lines = [
            f"---- {self.name} ----",
            f"--A: ({a.shape[0]}, {a.shape[1]})",
            f"data: {debug_list(a.data[:na])} .. ",
            f"indices: {debug_list(a.indices[:min(len(a.indices), 5)])} ...",
            f"indptr: {debug_list(a.indptr[:min(len(a.indptr), 5)])} ...",
            f"--B: ({b.shape[0]}, {b.shape[1]})",
            f"data: {debug_list(a.data[:nb_d])} ...",        # sic: A's data (gemm.rs:79)
            f"indices: {debug_list(a.indices[:nb_i])} ...",  # sic: A's indices (gemm.rs:84)
            f"indptr: {debug_list(b.indptr[:min(len(b.indptr), 5)])} ...",
        ]
        return "\n".join(lines) + "\n"
