"""Host CSR value types -- mirror of the functional part of the reference's src/storage.rs.

Only what the hot path's boundary needs: ``Element`` (storage.rs:22-32), ``CsrRow``
(storage.rs:34-126, note the reference's naming trap: ``CsrRow.indptr`` holds *column ids*
and ``rowptr`` the row id) and ``CsrMatStorage`` (storage.rs:150-324) with ``usize`` index
arrays, i.e. exactly the buffers the C ABI's ``spada_csr_view`` borrows.  The fiber cache,
psum storage and all access counters are the timing model and are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .gemm import GEMM
from .rustfmt import debug_list


@dataclass
class Element:
    idx: List[int]
    value: float


@dataclass
class CsrRow:
    rowptr: int
    data: np.ndarray = field(default_factory=lambda: np.empty(0, dtype=np.float64))
    indptr: np.ndarray = field(default_factory=lambda: np.empty(0, dtype=np.uint64))

    @classmethod
    def new_from_data(cls, rowptr: int, data, indptr) -> "CsrRow":
        """storage.rs:52-59."""
        return cls(rowptr, np.asarray(data, dtype=np.float64), np.asarray(indptr, dtype=np.uint64))

    def len(self) -> int:
        return len(self.indptr)

    def size(self) -> int:
        return len(self.data) + len(self.indptr)

    def as_element_vec(self) -> List[Element]:
        return [Element([self.rowptr, int(c)], float(d)) for d, c in zip(self.data, self.indptr)]

    def __str__(self) -> str:
        """storage.rs:115-125."""
        n = min(len(self.data), 5)
        return f"rowptr: {self.rowptr} indptr: {debug_list(self.indptr[:n])} data: {debug_list(self.data[:n])}"


class CsrMatStorage:
    """storage.rs:150-160: data Vec<f64>, indptr Vec<usize>, indices Vec<usize>, mat_shape [cols, rows]."""

    def __init__(self, data, indptr, indices, mat_shape):
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint64)
        self.mat_shape = list(mat_shape)  # [cols, rows] (sic, storage.rs:225, 236)
        self.read_count = 0
        self.write_count = 0
        self.remapped = False
        self.row_remap: Dict[int, int] = {}

    @classmethod
    def init_with_gemm(cls, gemm: GEMM):
        """storage.rs:214-239."""
        a, b = gemm.a, gemm.b
        sa = cls(a.data, a.indptr, a.indices, [a.shape[1], a.shape[0]])
        if b is a:
            sb = cls.__new__(cls)
            sb.__dict__.update(sa.__dict__)
            sb.row_remap = {}
        else:
            sb = cls(b.data, b.indptr, b.indices, [b.shape[1], b.shape[0]])
        return sa, sb

    def row_num(self) -> int:
        return len(self.indptr) - 1

    def get_ele_num(self, row_s: int, row_t: int) -> int:
        return int(self.indptr[row_t] - self.indptr[row_s]) if not self.remapped else sum(
            int(self.indptr[self.row_remap[i] + 1] - self.indptr[self.row_remap[i]]) for i in range(row_s, row_t))

    def rowptr(self, rowid: int) -> int:
        return int(self.indptr[self.row_remap[rowid]] if self.remapped else self.indptr[rowid])

    def read_row(self, row: int) -> CsrRow:
        r = self.row_remap[row] if self.remapped else row
        s, e = int(self.indptr[r]), int(self.indptr[r + 1])
        return CsrRow.new_from_data(row, self.data[s:e].copy(), self.indices[s:e].copy())

    def reorder_row(self, rowmap: Dict[int, int]):
        """storage.rs:252-255 (the -p preprocessing; C is unchanged by it, simulator.rs:1039-1060)."""
        self.remapped = True
        self.row_remap = rowmap


def sort_by_length(amat: CsrMatStorage) -> Dict[int, int]:
    """preprocessing.rs:76-89: stable sort of row ids by stored length."""
    lens = np.diff(amat.indptr.astype(np.int64))
    order = np.argsort(lens, kind="stable")
    return {i: int(r) for i, r in enumerate(order)}
