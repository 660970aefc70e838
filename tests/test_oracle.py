"""Pins the CPU oracle (oracle/spgemm_oracle.c).

The reference ships no tests or golden vectors ("parity unpinned"), so the
restatement is pinned against scipy's SMMP `A @ B` (+ sort_indices) -- bit-identical structure
and f64 bits -- and against the survey's known answers for the reference's only shipped
operand, matrices/cari.mtx (fixtures under tests/golden/, generator committed beside them).
"""
import hashlib
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, random_csr


def scipy_spgemm(a, b):
    c = (a @ b).tocsr()
    c.sort_indices()
    return c


def assert_bit_equal(cp, cj, cx, ref):
    assert np.array_equal(cp, ref.indptr.astype(np.int64))
    assert np.array_equal(cj, ref.indices.astype(np.int32))
    assert np.array_equal(cx.view(np.uint64), ref.data.astype(np.float64).view(np.uint64))


def test_cari_known_answers(oracle, cari):
    known = json.load(open(os.path.join(GOLDEN, "cari_known_answers.json")))
    a = cari
    assert list(a.shape) == known["a_shape"] and a.nnz == known["a_nnz"]
    b = oracle.transpose(a)  # gemm.rs:41-53: 400 x 1200 is not square => B = A^T
    bt = a.T.tocsr(); bt.sort_indices()
    assert np.array_equal(b.indptr, bt.indptr) and np.array_equal(b.indices, bt.indices)
    assert np.array_equal(b.data, bt.data)
    assert int(oracle.flops(a, b).sum()) == known["products"] == 57760800
    cp, cj, cx = oracle.spgemm(a, b)
    sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    assert len(cj) == known["c_nnz"] == 160000
    assert np.array_equal(cp, 400 * np.arange(401))
    assert sha(cp.astype("<i8")) == known["sha256_indptr_i64"]
    assert sha(cj.astype("<i8")) == known["sha256_indices_i64"]
    assert sha(cx.astype("<f8")) == known["sha256_data_f64"]
    assert float(cx.sum()) == known["c_data_sum"] == 7833.707235839987
    assert cj[:5].tolist() == known["row0_cols"]
    assert cx[:5].tolist() == known["row0_vals"]


def test_cari_vs_scipy_bits(oracle, cari):
    b = cari.T.tocsr(); b.sort_indices()
    assert_bit_equal(*oracle.spgemm(cari, b), scipy_spgemm(cari, b))


@pytest.mark.parametrize("m,k,n,density,seed", [(50, 40, 60, 0.1, 1), (200, 300, 100, 0.02, 2), (1, 1, 1, 1.0, 3),
                                                (64, 64, 64, 0.5, 4), (300, 10, 300, 0.3, 5)])
def test_random_vs_scipy_bits(oracle, m, k, n, density, seed):
    a = random_csr(m, k, density=density, seed=seed, values="signed")
    b = random_csr(k, n, density=density, seed=seed + 100, values="signed")
    assert_bit_equal(*oracle.spgemm(a, b), scipy_spgemm(a, b))


def test_threads_match_sequential(oracle):
    a = random_csr(500, 400, density=0.03, seed=7)
    b = random_csr(400, 450, density=0.03, seed=8)
    s = oracle.spgemm(a, b, threads=1)
    t = oracle.spgemm(a, b, threads=4)
    for x, y in zip(s, t):
        assert np.array_equal(x, y)


def test_structural_zeros_kept(oracle):
    # explicit zero in A, and products cancelling to 0.0 stay stored (simulator.rs:209-221)
    a = sp.csr_matrix((np.array([0.0, 1.0, 1.0]), np.array([0, 0, 1]), np.array([0, 1, 3])), shape=(2, 2))
    b = sp.csr_matrix((np.array([2.0, 3.0, -3.0]), np.array([0, 1, 1]), np.array([0, 2, 3])), shape=(2, 2))
    cp, cj, cx = oracle.spgemm(a, b)
    assert cp.tolist() == [0, 2, 4]
    assert cj.tolist() == [0, 1, 0, 1]
    assert cx.tolist() == [0.0, 0.0, 2.0, 0.0]


def test_empty_rows_and_empty_b_rows(oracle):
    # A row 0 empty; A row 1 only references an empty B row => both C rows empty (simulator.rs:1037-1044)
    a = sp.csr_matrix((np.array([1.0, 2.0]), np.array([1, 0]), np.array([0, 0, 1, 2])), shape=(3, 2))
    b = sp.csr_matrix((np.array([5.0]), np.array([2]), np.array([0, 1, 1])), shape=(2, 4))
    cp, cj, cx = oracle.spgemm(a, b)
    assert cp.tolist() == [0, 0, 0, 1] and cj.tolist() == [2] and cx.tolist() == [10.0]


def test_nan_inf_propagate(oracle):
    a = sp.csr_matrix((np.array([np.nan, np.inf]), np.array([0, 1]), np.array([0, 1, 2])), shape=(2, 2))
    b = sp.csr_matrix((np.array([1.0, -1.0]), np.array([0, 0]), np.array([0, 1, 2])), shape=(2, 1))
    _, _, cx = oracle.spgemm(a, b)
    assert np.isnan(cx[0]) and cx[1] == -np.inf


def test_no_fma_contraction(oracle):
    # fl(fl(a*b) + fl(c*d)) differs from an FMA-contracted evaluation for these operands
    a1, b1 = 1.0 + 2.0 ** -30, 1.0 + 2.0 ** -30
    a = sp.csr_matrix((np.array([a1, 1.0]), np.array([0, 1]), np.array([0, 2])), shape=(1, 2))
    b = sp.csr_matrix((np.array([b1, -1.0]), np.array([0, 0]), np.array([0, 1, 2])), shape=(2, 1))
    _, _, cx = oracle.spgemm(a, b)
    assert cx[0] == np.float64(a1 * b1) + np.float64(-1.0)


def test_flops_and_validate_and_groups(oracle):
    a = random_csr(100, 80, density=0.05, seed=11)
    b = random_csr(80, 90, density=0.05, seed=12)
    f = oracle.flops(a, b)
    assert np.array_equal(f, np.asarray((a != 0).astype(np.int64) @ np.diff(b.indptr)).ravel())
    assert oracle.validate_csr(a) == 0
    bad = a.copy()
    r = int(np.argmax(np.diff(a.indptr) >= 2))
    s = a.indptr[r]
    bad.indices[s], bad.indices[s + 1] = bad.indices[s + 1], bad.indices[s]
    assert oracle.validate_csr(bad) == r + 1
    # parse_group (rowwise_perf_adjust.rs:36-77): cari has 400 rows of exactly 382 nnz => one group
    lens = np.array([4, 5, 6, 0, 10, 11, 30, 2, 2])
    m = sp.csr_matrix((np.ones(lens.sum()), np.concatenate([np.arange(l) for l in lens]),
                       np.concatenate([[0], np.cumsum(lens)])), shape=(len(lens), 40))
    assert oracle.parse_group(m).tolist() == [0, 4, 6, 7]


def test_cari_single_group(oracle, cari):
    assert oracle.parse_group(cari).tolist() == [0]
