"""Seeded synthetic operands for the BASELINE.json configurations.

Every generator follows the survey's "generator rule": numpy ``default_rng(seed)``,
build COO -> ``tocsr()`` -> ``sum_duplicates()`` -> ``sort_indices()``, i.e. the same
canonical CSR the reference's loaders hand to the simulator (py2rust.rs:62-97:
``scipy.io.mmread(..).tocsr()``).  ``make_gemm`` then applies gemm.rs:41-53
(square => A x A, otherwise A x A^T).

The exact nnz counts quoted in SURVEY.md 8d act as generator known-answers
(``KNOWN`` below); bench.py asserts them at full size.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# full-size known answers: name -> (m, k, nnzA, products, nnzC or None)
KNOWN = {
    "poisson": (4194304, 4194304, 20963328, 104783880, 54484996),
    "er": (2097152, 2097152, 33554315, 536867278, 536836286),
    "rmat": (2097152, 2097152, 33541469, 5321533023, None),
    "rect": (1048576, 4194304, 32739036, 288281410, 251153314),
}


def _canon(m: sp.spmatrix) -> sp.csr_matrix:
    m = m.tocsr()
    m.sum_duplicates()
    m.sort_indices()
    m.data = m.data.astype(np.float64, copy=False)
    return m


def poisson2d(grid: int = 2048) -> sp.csr_matrix:
    """5-point stencil on a grid x grid mesh: kron(I,T)+kron(T,I), T=tridiag(-1,2,-1)."""
    t = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(grid, grid), format="csr")
    i = sp.identity(grid, format="csr")
    return _canon(sp.kron(i, t) + sp.kron(t, i))


def erdos_renyi(log2n: int = 21, per_row: int = 16, seed: int = 1234) -> sp.csr_matrix:
    """Exactly ``per_row`` draws per row, columns uniform, values U(0.001,1), dups summed."""
    n = 1 << log2n
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n, size=n * per_row)
    vals = rng.uniform(0.001, 1.0, size=n * per_row)
    rows = np.repeat(np.arange(n, dtype=np.int64), per_row)
    return _canon(sp.coo_matrix((vals, (rows, cols)), shape=(n, n)))


def rmat(scale: int = 21, edge_factor: int = 16, seed: int = 42,
         probs=(0.45, 0.22, 0.22, 0.11)) -> sp.csr_matrix:
    """R-MAT bit recursion, MSB first: row bit = u>=a+b; col bit = (a<=u<a+b) or (u>=a+b+c)."""
    a, b, c, _d = probs
    n = 1 << scale
    e = edge_factor * n
    rng = np.random.default_rng(seed)
    rows = np.zeros(e, dtype=np.int64)
    cols = np.zeros(e, dtype=np.int64)
    for _level in range(scale):
        u = rng.random(e)
        rbit = u >= a + b
        cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        rows = (rows << 1) | rbit
        cols = (cols << 1) | cbit
    vals = rng.uniform(0.001, 1.0, size=e)
    return _canon(sp.coo_matrix((vals, (rows, cols)), shape=(n, n)))


def rect_powerlaw(log2m: int = 20, log2k: int = 22, seed: int = 2024,
                  mean_target: float = 32.0, cap: int = 65536) -> sp.csr_matrix:
    """Rectangular power-law: row length clip(floor(x_m U^(-1/1.5)),1,cap), x_m=mean/3."""
    m, k = 1 << log2m, 1 << log2k
    rng = np.random.default_rng(seed)
    x_m = mean_target / 3.0
    u = rng.random(m)
    lens = np.clip(np.floor(x_m * u ** (-1.0 / 1.5)), 1, min(cap, k)).astype(np.int64)
    total = int(lens.sum())
    rows = np.repeat(np.arange(m, dtype=np.int64), lens)
    cols = rng.integers(0, k, size=total)
    vals = rng.uniform(0.001, 1.0, size=total)
    return _canon(sp.coo_matrix((vals, (rows, cols)), shape=(m, k)))


def make_gemm(a: sp.csr_matrix):
    """gemm.rs:41-53: B = A if square else A^T (as canonical CSR)."""
    if a.shape[0] == a.shape[1]:
        return a, a
    return a, _canon(a.T)


def build(name: str, scale: float = 1.0):
    """Return (A, B) for a BASELINE config name; ``scale`` < 1 shrinks log-sizes for tests.

    scale is applied as a reduction of log2 sizes: scale=1 -> full size, otherwise
    ``shift = round(-log2(scale))`` bits are removed from each dimension.
    """
    shift = 0 if scale >= 1.0 else int(round(-np.log2(scale)))
    if name == "poisson":
        a = poisson2d(max(4, 2048 >> shift))
    elif name == "er":
        a = erdos_renyi(max(6, 21 - shift))
    elif name == "rmat":
        a = rmat(max(6, 21 - shift))
    elif name == "rect":
        a = rect_powerlaw(max(6, 20 - shift), max(8, 22 - shift),
                          cap=max(64, 65536 >> shift))
    else:
        raise ValueError(f"unknown workload {name!r}")
    return make_gemm(a)


def algorithmic_bytes(nnz_a: int, m: int, nnz_b: int, k: int, nnz_c: int) -> int:
    """Compulsory traffic in the device layout: i32 col, f64 val, i64 row_ptr."""
    return (12 * nnz_a + 8 * (m + 1)) + (12 * nnz_b + 8 * (k + 1)) + (12 * nnz_c + 8 * (m + 1))
