"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- needs a B200.

Bar: row_ptr and col_idx bit-exact; values within relative 1e-12 per
entry in f64 (TOL below).  Every path of the engine sums the products of one C[i,j] in the
oracle's order (ascending k, left to right) -- the sort bins through the arrival index in the sort
key, the long rows (> 4096 products) through stable merges of unreduced chunks -- so values are
checked BIT-EXACT everywhere, with signed (cancelling) operands.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, random_csr

pytestmark = pytest.mark.gpu

TOL = 1e-12  # relative, per entry (north_star)


def check(result, ref, exact_values):
    ip, ix, dx = result.to_host()
    cp, cj, cx = ref
    assert ip.dtype == np.int64 and ix.dtype == np.int32 and dx.dtype == np.float64
    assert np.array_equal(ip, cp), "row_ptr differs"
    assert np.array_equal(ix, cj), "col_idx differs"
    if exact_values:
        same = (dx.view(np.uint64) == cx.view(np.uint64)) | (np.isnan(dx) & np.isnan(cx))
        assert same.all(), f"{(~same).sum()} values differ in bits"
    else:
        finite = np.isfinite(cx)
        assert np.array_equal(np.isnan(dx), np.isnan(cx))
        assert np.array_equal(dx[~finite & ~np.isnan(cx)], cx[~finite & ~np.isnan(cx)])
        err = np.abs(dx[finite] - cx[finite])
        assert (err <= TOL * np.abs(cx[finite])).all(), f"max rel err {np.max(err / np.maximum(np.abs(cx[finite]), 1e-300))}"


def run(engine, oracle, a, b, exact=True, usize=False):
    r = engine.spgemm(a, b, usize=usize)
    ref = oracle.spgemm(a, b, threads=oracle.max_threads())
    check(r, ref, exact)
    st = r.stats()
    assert st["nnz_c"] == len(ref[1])
    assert st["products"] == int(oracle.flops(a, b).sum())
    return r, st


# ---- every bin, forced individually ---------------------------------------------------------
@pytest.mark.parametrize("ka,lb,expect_bin", [
    (1, 1, "32"), (4, 8, "32"), (5, 5, "32"), (8, 8, "64"), (10, 12, "128"), (16, 16, "256"),
    (20, 25, "512"), (32, 32, "1024"), (40, 50, "2048"), (64, 64, "4096"), (70, 100, "8192"), (128, 64, "8192"),
    (300, 40, "16384"), (300, 230, "131072"), (517, 300, "262144"),
])
def test_each_bin(engine, oracle, ka, lb, expect_bin):
    m, k, n = 257, 600, 5000
    a = random_csr(m, k, row_nnz=ka, seed=ka * 7 + lb, values="signed")
    b = random_csr(k, n, row_nnz=lb, seed=ka * 11 + lb, values="signed")
    _, st = run(engine, oracle, a, b)
    assert list(st["bins"].keys()) == [expect_bin], st["bins"]
    assert st["bins"][expect_bin]["rows"] == m


@pytest.mark.parametrize("n_cols", [1 << 10, 1 << 20, 1 << 21, 1 << 22, 1 << 23, (1 << 28) + 5])
def test_key_width_paths(engine, oracle, n_cols):
    # 32-bit packed (column, arrival) keys when they fit; one bit short: sorted without the top bit of the column, then
    # split by it; 64-bit keys otherwise
    for ka, lb in [(4, 6), (16, 16), (30, 30), (40, 50), (60, 60), (80, 90), (200, 120)]:
        a = random_csr(130, 300, row_nnz=ka, seed=3, values="signed")
        b = random_csr(300, n_cols, row_nnz=lb, seed=4, values="signed")
        run(engine, oracle, a, b)


@pytest.mark.parametrize("n_cols,ka,lb", [(1 << 23, 30, 30), (1 << 22, 40, 50), (1 << 21, 60, 60), (1 << 21, 80, 90),
                                          (1 << 21, 300, 230), ((1 << 21) - 3, 64, 64)])
def test_top_bit_split_with_sums(engine, oracle, n_cols, ka, lb):
    # the one-bit-short path on every sort width (1024 / 2048 / 4096 / chunks of long rows): B's columns come from a
    # small pool on both sides of the bit the key drops, so runs of equal columns exist in both halves of the split,
    # and columns that differ only in that bit must not be summed together
    rng = np.random.default_rng(n_cols % 1000 + ka)
    half = n_cols // 2 if n_cols & (n_cols - 1) == 0 else 1 << (n_cols.bit_length() - 1)
    pool = np.unique(np.concatenate([rng.choice(700, 400, replace=False), half + rng.choice(700, 400, replace=False),
                                     [0, half - 1, half, n_cols - 1]]))
    pool = pool[pool < n_cols]
    k = 300
    rows = np.repeat(np.arange(k), lb)
    cols = np.concatenate([rng.choice(pool, lb, replace=False) for _ in range(k)])
    b = sp.csr_matrix((rng.uniform(-1, 1, len(rows)), (rows, cols)), shape=(k, n_cols))
    b.sort_indices()
    a = random_csr(70, k, row_nnz=ka, seed=5, values="signed")
    run(engine, oracle, a, b)
    b_low = b[:, :half].tocsr()   # no top bits at all / only top bits
    run(engine, oracle, a, sp.csr_matrix((b_low.data, b_low.indices, b_low.indptr), shape=(k, n_cols)))


def test_mixed_bins_and_permutation(engine, oracle):
    rng = np.random.default_rng(5)
    lens = rng.choice([0, 1, 3, 8, 20, 60, 150, 400], size=3000, p=[.1, .2, .2, .2, .15, .1, .04, .01])
    a = random_csr(3000, 2000, row_nnz=lens, seed=6, values="signed")
    b = random_csr(2000, 4000, row_nnz=rng.choice([0, 2, 9, 30], size=2000), seed=7, values="signed")
    _, st = run(engine, oracle, a, b)
    assert len(st["bins"]) >= 5


# ---- the reference's own operand ------------------------------------------------------------
def test_cari_golden(engine, oracle, cari, spada):
    known = json.load(open(os.path.join(GOLDEN, "cari_known_answers.json")))
    g = spada.GEMM.from_mat("cari", cari)
    assert g.b.shape == (1200, 400)
    r, st = run(engine, oracle, g.a, g.b)  # 144k products per row -> long rows, 6 merge levels; bit-exact
    ip, ix, dx = r.to_host()
    sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    assert sha(ip.astype("<i8")) == known["sha256_indptr_i64"]
    assert sha(ix.astype("<i8")) == known["sha256_indices_i64"]
    assert st["products"] == known["products"] and st["nnz_c"] == known["c_nnz"]
    assert abs(dx.sum() - known["c_data_sum"]) <= 1e-9
    assert dx[:5].tolist() == known["row0_vals"]
    assert dx.sum() == known["c_data_sum"]
    assert sha(dx.astype("<f8")) == known["sha256_data_f64"]


# ---- edge cases of SURVEY.md 8a ---------------------------------------------------------------
def test_empty_rows_empty_b_rows_explicit_zeros(engine, oracle):
    a = sp.csr_matrix((np.array([0.0, 1.0, 1.0, 2.0]), np.array([0, 0, 1, 2]), np.array([0, 0, 1, 3, 4])), shape=(4, 3))
    b = sp.csr_matrix((np.array([2.0, 3.0, -3.0]), np.array([0, 1, 1]), np.array([0, 2, 3, 3])), shape=(3, 2))
    r, _ = run(engine, oracle, a, b)
    ip, ix, dx = r.to_host()
    assert ip.tolist() == [0, 0, 2, 4, 4]          # row 0 empty; row 3 only hits the empty B row
    assert dx.tolist() == [0.0, 0.0, 2.0, 0.0]     # explicit / cancelled zeros stay structural


def test_zero_sized(engine, oracle, spada):
    a = sp.csr_matrix((0, 5)); b = sp.csr_matrix((5, 7))
    r = engine.spgemm(a, b)
    assert r.shape == (0, 7) and r.nnz == 0 and r.to_host()[0].tolist() == [0]
    a = sp.csr_matrix((6, 5)); r = engine.spgemm(a, b)
    assert r.nnz == 0 and r.to_host()[0].tolist() == [0] * 7
    a = sp.csr_matrix((3, 0)); b = sp.csr_matrix((0, 4))
    assert engine.spgemm(a, b).nnz == 0


def test_nan_inf_propagate(engine, oracle):
    a = sp.csr_matrix((np.array([np.nan, np.inf, 1.0]), np.array([0, 1, 1]), np.array([0, 1, 2, 3])), shape=(3, 2))
    b = sp.csr_matrix((np.array([1.0, -1.0, 0.0]), np.array([0, 0, 1]), np.array([0, 1, 3])), shape=(2, 2))
    run(engine, oracle, a, b)


def test_no_fma(engine, oracle):
    a1 = 1.0 + 2.0 ** -30
    a = sp.csr_matrix((np.array([a1, 1.0]), np.array([0, 1]), np.array([0, 2])), shape=(1, 2))
    b = sp.csr_matrix((np.array([a1, -1.0]), np.array([0, 0]), np.array([0, 1, 2])), shape=(2, 1))
    r, _ = run(engine, oracle, a, b)
    assert r.to_host()[2][0] == np.float64(a1 * a1) + np.float64(-1.0)


def test_errors(engine, spada):
    a = random_csr(10, 7, density=0.3, seed=1)
    b = random_csr(8, 5, density=0.3, seed=2)
    with pytest.raises(spada.SpadaB200Error) as e:
        engine.spgemm(a, b)
    assert e.value.status == "DIM_MISMATCH"
    bad = random_csr(10, 8, row_nnz=4, seed=3)
    s = bad.indptr[2]
    bad.indices[s], bad.indices[s + 1] = bad.indices[s + 1], bad.indices[s]
    for usize in (False, True):
        with pytest.raises(spada.SpadaB200Error) as e:
            engine.spgemm(bad, b, usize=usize)
        assert e.value.status == "UNSORTED_INPUT"
    dup = random_csr(10, 8, row_nnz=4, seed=3)
    dup.indices[dup.indptr[5] + 1] = dup.indices[dup.indptr[5]]
    with pytest.raises(spada.SpadaB200Error) as e:
        engine.upload(dup)
    assert e.value.status == "UNSORTED_INPUT"
    oob = random_csr(10, 8, row_nnz=2, seed=4)
    oob.indices[-1] = 8
    with pytest.raises(spada.SpadaB200Error):
        engine.upload(oob)


def test_long_rows_and_first_last(engine, oracle):
    # one A row far longer than 256 nonzeros (CTA-per-row flop counter), first and last rows non-trivial
    lens = np.full(64, 3); lens[0] = 1500; lens[-1] = 700; lens[10] = 0
    a = random_csr(64, 4000, row_nnz=lens, seed=8)
    b = random_csr(4000, 30000, row_nnz=np.random.default_rng(9).integers(0, 12, size=4000), seed=10)
    run(engine, oracle, a, b)


def test_many_empty_b_rows_in_a_row(engine, oracle):
    # an A row with > 32 nonzeros whose B rows are mostly empty: products <= 32 but several expansion batches
    a = random_csr(40, 500, row_nnz=100, seed=11)
    lb = np.zeros(500, dtype=np.int64); lb[::37] = 2
    b = random_csr(500, 64, row_nnz=lb, seed=12)
    _, st = run(engine, oracle, a, b)
    assert set(st["bins"]) <= {"empty", "32"}


def test_high_compression_row(engine, oracle):
    # every product of a row lands on few columns (long equal-column runs in the segmented sum)
    a = random_csr(50, 64, row_nnz=60, seed=13, values="signed")
    b = random_csr(64, 8, row_nnz=8, seed=14, values="signed")
    run(engine, oracle, a, b)
    b2 = random_csr(64, 40, row_nnz=40, seed=15, values="signed")   # 2400 products -> 40 outputs
    run(engine, oracle, a, b2)


def _widen(b, n_cols, shift=0):
    """Same entries, columns moved by `shift`, in a matrix that is n_cols wide."""
    return sp.csr_matrix((b.data, b.indices + shift, b.indptr), shape=(b.shape[0], n_cols))


@pytest.mark.parametrize("ka,lb,width", [(30, 30, 64), (50, 40, 300), (100, 90, 700), (200, 60, 2000), (90, 80, 1 << 14)])
def test_skewed_columns(engine, oracle, ka, lb, width):
    # every product of a row lands in a narrow band of a 2^20-wide B: long runs of equal columns in the sort bins,
    # many equal keys per merge tile in the long rows
    a = random_csr(120, 400, row_nnz=ka, seed=ka + 1)
    b = _widen(random_csr(400, width, row_nnz=min(lb, width), seed=lb + 2), 1 << 20, shift=(1 << 19) + 77)
    run(engine, oracle, a, b)


def test_long_rows_mixed_lengths(engine, oracle):
    # rows of 513 .. ~40000 products side by side: CTA-per-row sort bins and long rows in one call
    rng = np.random.default_rng(21)
    lens = rng.choice([0, 30, 70, 130, 260, 520, 900, 2500], size=400, p=[.05, .2, .2, .2, .15, .1, .07, .03])
    a = random_csr(400, 6000, row_nnz=lens, seed=22, values="signed")
    b = random_csr(6000, 1 << 20, row_nnz=rng.integers(10, 26, size=6000), seed=23, values="signed")
    run(engine, oracle, a, b)
    # tiny column space: every output column is hit hundreds of times (long runs of equal columns across chunks)
    b2 = random_csr(6000, 48, row_nnz=rng.integers(8, 20, size=6000), seed=24, values="signed")
    run(engine, oracle, a, b2)


@pytest.mark.parametrize("pad", ["0", "1", "16"])
def test_fiber_store_layouts(spada, oracle, monkeypatch, pad):
    # B is gathered from its fiber store: no store / descriptors into the canonical arrays / rows on 64-byte boundaries
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    monkeypatch.setenv("SPADA_B200_FIBER_PAD", pad)
    e = spada.Engine(two_phase=True)
    try:
        rng = np.random.default_rng(31)
        for ka, lb, n in [(3, 4, 900), (16, 16, 1 << 18), (40, 50, 5000), (300, 40, 5000), (300, 230, 5000)]:
            a = random_csr(150, 700, row_nnz=ka, seed=ka + 40)
            b = random_csr(700, n, row_nnz=rng.integers(0, 2 * lb, size=700), seed=lb + 41)
            # resident operands: B's fiber store is built on its first use (the host-level call skips it)
            check(e.spgemm_dev(e.upload(a), e.upload(b)), oracle.spgemm(a, b, threads=oracle.max_threads()), True)
        a = random_csr(400, 400, density=0.03, seed=42)     # A x A: one operand on both sides
        da = e.upload(a)
        check(e.spgemm_dev(da, da), oracle.spgemm(a, a, threads=oracle.max_threads()), True)
    finally:
        e.close()


def test_prepare_wrapped_operand(engine, oracle):
    a = random_csr(300, 500, row_nnz=12, seed=43)
    b = random_csr(500, 4000, row_nnz=np.random.default_rng(44).integers(0, 30, size=500), seed=45)
    da, db = engine.upload(a), engine.upload(b)
    p, c, v = db.device_ptrs()
    wb = engine.wrap_device(b.shape[0], b.shape[1], b.nnz, p, c, v, keepalive=db)
    ref = oracle.spgemm(a, b, threads=oracle.max_threads())
    check(engine.spgemm_dev(da, wb), ref, True)      # wrapped arrays: row_ptr path
    assert wb.prepare() >= 0.0
    check(engine.spgemm_dev(da, wb), ref, True)      # same arrays through the fiber store


@pytest.mark.parametrize("shape,row_nnz", [((1, 1), 1), ((400, 1200), 30), ((3000, 70000), None), ((5000, 257), 12),
                                            ((64, 1 << 22), 3000), ((2000, 300), 0)])
def test_device_transpose_matches_scipy(engine, shape, row_nnz):
    # gemm.rs:41-53: B = A^T as CSR for non-square SS workloads -- values moved, structure bit-identical to scipy
    rng = np.random.default_rng(shape[0] + shape[1])
    lens = rng.integers(0, 40, size=shape[0]) if row_nnz is None else row_nnz
    a = random_csr(shape[0], shape[1], row_nnz=lens, seed=51, values="signed")
    t = engine.transpose(engine.upload(a)).to_scipy()
    ref = a.T.tocsr(); ref.sort_indices()
    assert t.shape == ref.shape
    assert np.array_equal(t.indptr, ref.indptr) and np.array_equal(t.indices, ref.indices)
    assert np.array_equal(t.data.view(np.uint64), ref.data.view(np.uint64))


def test_transposed_operand_feeds_the_hot_path(engine, oracle, cari, spada):
    g = spada.GEMM.from_mat("cari", cari)                 # host transpose, as the reference does it
    da = engine.upload(g.a)
    db = engine.transpose(da)                             # device transpose
    check(engine.spgemm_dev(da, db), oracle.spgemm(g.a, g.b, threads=oracle.max_threads()), True)


def test_long_rows_in_waves(spada, oracle, monkeypatch):
    # a 1 MiB ping-pong budget forces the long rows through many waves (level bins cut by their per-row bound)
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(60)
    lens = rng.choice([10, 40, 90, 300, 600], size=300)
    a = random_csr(300, 800, row_nnz=lens, seed=61, values="signed")
    b = random_csr(800, 200000, row_nnz=rng.integers(100, 260, size=800), seed=62, values="signed")
    ref = oracle.spgemm(a, b, threads=oracle.max_threads())
    for mb in ("1", "64"):
        monkeypatch.setenv("SPADA_B200_LONG_WS_MB", mb)
        for serial in (False, True):
            e = spada.Engine(two_phase=True, serial=serial)
            try:
                r = e.spgemm(a, b)
                assert set(r.stats()["bins"]) <= set(spada.engine.LONG_BINS) | {"4096", "2048", "1024"}
                check(r, ref, True)
            finally:
                e.close()


def test_long_row_shapes(engine, oracle):
    # the shapes a chunked row can take: one B row longer than several chunks, a row of exactly k * 4096 products,
    # a row one product over, A entries that meet empty B rows between the chunks, and duplicate-heavy columns
    k, n = 64, 50000
    lb = np.zeros(k, dtype=np.int64)
    lb[0] = 20000; lb[1] = 4096; lb[2] = 4096; lb[3] = 1; lb[5] = 8191; lb[7] = 3000; lb[9:20] = 700
    b = random_csr(k, n, row_nnz=lb, seed=70, values="signed")
    rows = [[0], [1, 2], [1, 2, 3], [0, 4, 5, 6, 7], list(range(64)), [4, 6, 8], [5, 8], list(range(9, 20)), [0, 5]]
    ip = np.cumsum([0] + [len(r) for r in rows])
    ix = np.concatenate([np.array(r) for r in rows])
    a = sp.csr_matrix((np.random.default_rng(71).uniform(-1, 1, len(ix)), ix, ip), shape=(len(rows), k))
    run(engine, oracle, a, b)
    # few columns: every chunk carries every column, the runs of equal columns span all chunks
    b2 = random_csr(k, 30, row_nnz=np.minimum(lb, 30), seed=72, values="signed")
    a2 = random_csr(40, k, row_nnz=60, seed=73, values="signed")
    b2 = sp.vstack([b2] * 1).tocsr()
    a3 = sp.hstack([a2] * 6).tocsr(); a3.sort_indices()
    b3 = sp.vstack([b2] * 6).tocsr(); b3.sort_indices()
    run(engine, oracle, a3, b3)


def test_usize_layout_matches(engine, oracle):
    a = random_csr(300, 200, density=0.05, seed=16)
    b = random_csr(200, 250, density=0.05, seed=17)
    r32 = engine.spgemm(a, b)
    r64 = engine.spgemm(a, b, usize=True)
    ip, ix, dx = r64.to_host_usize()
    jp, jx, ex = r32.to_host()
    assert ip.dtype == np.uint64 and ix.dtype == np.uint64
    assert np.array_equal(ip.astype(np.int64), jp) and np.array_equal(ix.astype(np.int32), jx)
    assert np.array_equal(dx, ex)


def test_square_alias_a_times_a(engine, oracle):
    a = random_csr(500, 500, density=0.02, seed=18)
    run(engine, oracle, a, a)


def test_row_shards_concatenate(engine, oracle):
    a = random_csr(1000, 800, density=0.01, seed=19)
    b = random_csr(800, 900, density=0.01, seed=20)
    da, db = engine.upload(a), engine.upload(b)
    full = engine.spgemm_dev(da, db).to_host()
    bounds = engine.plan_shards(da, db, 4)
    assert bounds[0] == 0 and bounds[-1] == 1000 and np.all(np.diff(bounds) >= 0)
    total, f = engine.flops(da, db, per_row=True)
    assert np.array_equal(f.astype(np.int64), oracle.flops(a, b)) and total == f.sum()
    w = f.astype(np.int64) + 1
    shard_w = [w[bounds[i]:bounds[i + 1]].sum() for i in range(4)]
    assert max(shard_w) - min(shard_w) <= 2 * w.max()
    parts = [engine.spgemm_dev(da, db, int(bounds[i]), int(bounds[i + 1])).to_host() for i in range(4)]
    ip = np.concatenate([[0]] + [p[0][1:] + off for p, off in
                                 zip(parts, np.cumsum([0] + [p[0][-1] for p in parts[:-1]]))])
    assert np.array_equal(ip, full[0])
    assert np.array_equal(np.concatenate([p[1] for p in parts]), full[1])
    assert np.array_equal(np.concatenate([p[2] for p in parts]), full[2])


# ---- BASELINE configs, scaled down, against the oracle ------------------------------------------
@pytest.mark.parametrize("name,scale,exact", [("poisson", 1 / 16, True), ("er", 1 / 64, True),
                                              ("rmat", 1 / 128, True), ("rect", 1 / 64, True)])
def test_configs_scaled(engine, oracle, spada, name, scale, exact):
    a, b = spada.workloads.build(name, scale)
    run(engine, oracle, a, b, exact=exact)
    # the same product with device-resident operands (what bench.py times): B gathered through its fiber store
    da = engine.upload(a)
    db = da if b is a else engine.upload(b)
    check(engine.spgemm_dev(da, db), oracle.spgemm(a, b, threads=oracle.max_threads()), exact)


# ---- full size: properties that need no CPU product ------------------------------------------------
def _linearity(a, b, ip, ix, dx):
    # C 1 = A (B 1), row by row
    lhs = np.add.reduceat(np.append(dx, 0.0), ip[:-1])
    lhs[np.diff(ip) == 0] = 0.0
    rhs = a @ np.asarray(b.sum(axis=1)).ravel()
    assert np.allclose(lhs, rhs, rtol=1e-9, atol=1e-12)


def test_poisson_full_size(engine, spada):
    a, b = spada.workloads.build("poisson")
    m, k, nnz_a, products, nnz_c = spada.workloads.KNOWN["poisson"]
    r = engine.spgemm(a, b)
    st = r.stats()
    assert st["products"] == products and st["nnz_c"] == nnz_c
    ip, ix, dx = r.to_host()
    ref = (a @ b).tocsr(); ref.sort_indices()   # exact small integers: any association is bit-identical
    assert np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices) and np.array_equal(dx, ref.data)


def test_rect_full_size_vs_oracle(engine, oracle, spada):
    # BASELINE configs[4] at full size against the oracle, bit for bit (288 M products; the oracle takes ~1.5 s on 16 cores)
    a, b = spada.workloads.build("rect")
    da = engine.upload(a)
    db = engine.transpose(da)
    r = engine.spgemm_dev(da, db)
    check(r, oracle.spgemm(a, b, threads=oracle.max_threads()), True)
    assert r.stats()["nnz_c"] == spada.workloads.KNOWN["rect"][4]


def test_er_full_size_vs_oracle(engine, oracle, spada):
    # BASELINE configs[2] at full size against the oracle, bit for bit (537 M products, 6.4 GB of C)
    a, b = spada.workloads.build("er")
    da = engine.upload(a)
    r = engine.spgemm_dev(da, da)
    assert r.stats()["nnz_c"] == spada.workloads.KNOWN["er"][4]
    check(r, oracle.spgemm(a, b, threads=oracle.max_threads()), True)


def test_rmat_full_size_row_sample(engine, oracle, spada):
    # BASELINE configs[3]: the 200 heaviest rows (up to 1.1 M products each, 9 merge levels) and 10 000 random rows
    # of the full-size operand against the oracle, bit for bit; B is the whole 2M x 2M matrix
    a, b = spada.workloads.build("rmat")
    f = oracle.flops(a, b)
    assert int(f.sum()) == spada.workloads.KNOWN["rmat"][3]
    rows = np.union1d(np.argsort(f)[-200:], np.random.default_rng(7).choice(a.shape[0], 10000, replace=False))
    sub = a[rows]
    sub.sort_indices()
    db = engine.upload(b)
    r = engine.spgemm_dev(engine.upload(sub), db)
    st = r.stats()
    assert max(int(k) for k in st["bins"] if k != "empty") >= 1 << 20
    check(r, oracle.spgemm(sub, b, threads=oracle.max_threads()), True)


def test_signed_zero_survives(engine, oracle):
    # a column whose only product is -0.0 stays -0.0 (nothing is added to a +0.0 initial value), in the sort bins, the
    # merge path of the long rows and the dense accumulator path
    for ka, lb, n in [(3, 4, 500), (90, 80, 1 << 15), (90, 80, 300)]:
        a = random_csr(40, 200, row_nnz=ka, seed=94, values="signed")
        b = random_csr(200, n, row_nnz=min(lb, n), seed=95, values="signed")
        a.data[::7] = -0.0
        b.data[::5] = 0.0
        b.data = np.abs(b.data)
        a.data[a.indptr[3]:a.indptr[4]] = -0.0      # every product of row 3 is -0.0: so is every sum of that row
        r, _ = run(engine, oracle, a, b)
        ip, _, dx = r.to_host()
        row3 = dx[ip[3]:ip[4]]
        assert len(row3) and (row3 == 0.0).all() and np.signbit(row3).all()


def test_rect_full_size_properties(engine, spada):
    a, b = spada.workloads.build("rect")
    m, k, nnz_a, products, nnz_c = spada.workloads.KNOWN["rect"]
    assert a.nnz == nnz_a
    da = engine.upload(a)
    db = engine.transpose(da)              # B = A^T built on the device, resident operands as in bench.py
    tb = db.to_scipy()
    assert np.array_equal(tb.indptr, b.indptr) and np.array_equal(tb.indices, b.indices) and np.array_equal(tb.data, b.data)
    r = engine.spgemm_dev(da, db)
    st = r.stats()
    assert st["products"] == products and st["nnz_c"] == nnz_c
    ip, ix, dx = r.to_host()
    assert ip[0] == 0 and ip[-1] == nnz_c and np.all(np.diff(ip) >= 0)
    ok = np.diff(ix.astype(np.int64)) > 0
    inner = ip[1:-1]
    ok[inner[(inner > 0) & (inner < len(ix))] - 1] = True   # a descent is allowed only across a row boundary
    assert ok.all(), "columns not strictly ascending inside a row"
    assert ix.min() >= 0 and ix.max() < b.shape[1]
    _linearity(a, b, ip, ix, dx)
    # A x A^T is symmetric: spot-check C[i,j] == C[j,i] on sampled entries
    rng = np.random.default_rng(0)
    c = sp.csr_matrix((dx, ix, ip), shape=(m, m))
    rows = rng.integers(0, m, size=200)
    for i in rows:
        s, e = ip[i], ip[i + 1]
        if e > s:
            j = ix[s + (e - s) // 2]
            assert abs(c[i, j] - c[j, i]) <= 1e-12 * abs(c[i, j])


@pytest.mark.parametrize("acc", ["ip", "op", "multirow", "spada"])
def test_accelerator_policies_do_not_change_c(spada, oracle, acc):
    # frontend.rs:33-41 / scheduler.rs:729-753: the accelerator only picks the window shape
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    a, b = spada.workloads.build("er", 1 / 256)
    e = spada.Engine(accelerator=acc, block_shape=(4, 2))
    try:
        check(e.spgemm(a, b), oracle.spgemm(a, b, threads=oracle.max_threads()), True)
    finally:
        e.close()


# ---- sharded products, row-panel streaming ------------------------------------------------------------------------
def test_shards_into_one_buffer(engine, oracle, spada):
    # the two halves of a sharded product on ONE GPU: three shards stored one after the other into the same full-size
    # C buffers at their global offsets (host-known offsets and the device-side offset array) == the whole product
    rng = np.random.default_rng(80)
    lens = rng.choice([0, 2, 9, 40, 150, 500], size=900, p=[.1, .3, .3, .2, .07, .03])
    a = random_csr(900, 1500, row_nnz=lens, seed=81, values="signed")
    b = random_csr(1500, 30000, row_nnz=rng.integers(0, 60, size=1500), seed=82, values="signed")
    ref = oracle.spgemm(a, b, threads=oracle.max_threads())
    da, db = engine.upload(a), engine.upload(b)
    bounds = engine.plan_shards(da, db, 3)
    cap = engine.flops(da, db)
    for device_offsets in (False, True):
        buf = engine.cbuf_create(a.shape[0], b.shape[1], cap)
        nnz, off = [], 0
        arr = None
        if device_offsets:
            import torch
            arr = torch.zeros(3, dtype=torch.int64, device="cuda")
        for i in range(3):
            sh = engine.shard_begin(da, db, int(bounds[i]), int(bounds[i + 1]), host_nnz=True)   # one shard in flight per handle
            n_i = sh.nnz_local
            if device_offsets:
                arr[i] = n_i
                torch.cuda.synchronize()
                st = sh.finish([buf], 0, arr.data_ptr(), i)
            else:
                st = sh.finish([buf], off, 0, i)
            assert st["rows"] == bounds[i + 1] - bounds[i]
            off += n_i
            nnz.append(n_i)
        assert buf.nnz == len(ref[1]) == sum(nnz)
        ip, ix, dx = buf.to_host()
        assert np.array_equal(ip, ref[0]) and np.array_equal(ix, ref[1])
        assert np.array_equal(dx.view(np.uint64), ref[2].view(np.uint64))
        buf.free()


def test_group_of_all_gpus(spada, oracle):
    # every GPU of this process behind one call (what the single-process CLI drives); needs at least two devices
    n = spada.device_count()
    if n < 2:
        pytest.skip("needs two CUDA devices")
    rng = np.random.default_rng(83)
    lens = rng.choice([0, 3, 20, 90, 400], size=2000, p=[.1, .4, .3, .15, .05])
    a = random_csr(2000, 1800, row_nnz=lens, seed=84, values="signed")
    b = random_csr(1800, 50000, row_nnz=rng.integers(0, 80, size=1800), seed=85, values="signed")
    ref = oracle.spgemm(a, b, threads=oracle.max_threads())
    g = spada.Group(min(n, 8))
    try:
        for _ in range(2):
            check(g.spgemm(a, b), ref, True)
    finally:
        g.close()


@pytest.mark.parametrize("panel_products", [0, 3000, 200000])
def test_row_panel_streaming(engine, oracle, panel_products):
    # C in row panels handed to a sink (D2H of a panel beside the next panel's kernels) and straight into host arrays
    rng = np.random.default_rng(86)
    lens = rng.choice([0, 2, 9, 40, 150, 900], size=1200, p=[.1, .3, .3, .2, .07, .03])
    a = random_csr(1200, 1500, row_nnz=lens, seed=87, values="signed")
    b = random_csr(1500, 20000, row_nnz=rng.integers(0, 50, size=1500), seed=88, values="signed")
    ref = oracle.spgemm(a, b, threads=oracle.max_threads())
    da, db = engine.upload(a), engine.upload(b)
    got_p, got_c, got_v, spans = [np.zeros(1, dtype=np.int64)], [], [], []

    def sink(r0, r1, n0, ip, ix, dx):
        assert ip[0] == n0 and len(ip) == r1 - r0 + 1
        spans.append((r0, r1))
        got_p.append(ip[1:].copy()); got_c.append(ix.copy()); got_v.append(dx.copy())
    st = engine.spgemm_stream(da, db, sink, panel_products)
    assert spans[0][0] == 0 and spans[-1][1] == a.shape[0] and all(x[1] == y[0] for x, y in zip(spans, spans[1:]))
    assert st["panels"] == len(spans) and st["nnz_c"] == len(ref[1])
    if panel_products == 3000:
        assert st["panels"] > 20
    assert np.array_equal(np.concatenate(got_p), ref[0]) and np.array_equal(np.concatenate(got_c), ref[1])
    assert np.array_equal(np.concatenate(got_v).view(np.uint64), ref[2].view(np.uint64))
    ip = np.zeros(a.shape[0] + 1, dtype=np.int64); ix = np.zeros(len(ref[1]) + 7, dtype=np.int32); dx = np.zeros(len(ref[1]) + 7)
    st = engine.spgemm_to_host(da, db, ip, ix, dx, panel_products)
    n = st["nnz_c"]
    assert np.array_equal(ip, ref[0]) and np.array_equal(ix[:n], ref[1]) and np.array_equal(dx[:n].view(np.uint64), ref[2].view(np.uint64))
    with pytest.raises(Exception):
        engine.spgemm_to_host(da, db, ip, ix[:10], dx[:10], panel_products)     # capacity below nnz(C)


def test_sink_error_aborts(engine):
    a = random_csr(300, 300, density=0.02, seed=89)
    da = engine.upload(a)

    def bad(*_):
        raise ValueError("sink failed")
    with pytest.raises(ValueError):
        engine.spgemm_stream(da, da, bad, 500)
    check(engine.spgemm_dev(da, da), (lambda c: (c.indptr.astype(np.int64), c.indices, c.data))(
        (lambda c: (c.sort_indices(), c)[1])((a @ a).tocsr())), False)   # the engine is still usable


# ---- window choice (a-6): picked per operand / accelerator, reported, never changes C ---------------------------------
def test_window_choice_follows_the_operand(spada, oracle):
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    b = random_csr(400, 3000, row_nnz=2, seed=91, values="signed")
    short = random_csr(500, 400, row_nnz=5, seed=92, values="signed")     # 10 products from 5 A entries: fits [4, 8]
    wide = random_csr(500, 400, row_nnz=14, seed=93, values="signed")     # 28 products from 14 A entries: does not
    cases = [
        (dict(accelerator="spada"), short, [4, 8]), (dict(accelerator="spada"), wide, [1, 32]),
        (dict(accelerator="ip"), short, [1, 32]), (dict(accelerator="op"), short, [4, 8]),
        (dict(accelerator="op", lane_num=32), short, [1, 32]),
        (dict(accelerator="multirow", block_shape=(4, 2)), short, [4, 8]),
        (dict(accelerator="multirow", block_shape=(2, 4)), short, [1, 32]),
    ]
    for kw, a, want in cases:
        e = spada.Engine(single_pass=True, **kw)      # the single pass is where bin 1 has two window shapes
        try:
            r = e.spgemm(a, b)
            st = r.stats()
            assert list(st["bins"]) == ["32"], st["bins"]
            assert st["bins"]["32"]["window"] == want, (kw, st["bins"])
            check(r, oracle.spgemm(a, b, threads=oracle.max_threads()), True)
        finally:
            e.close()


@pytest.mark.parametrize("tile_pass", ["1", "0"])
def test_single_pass_on_mixed_rows(spada, oracle, monkeypatch, tile_pass):
    # mixed row lengths in one pass: tiles cut by work (SPADA_B200_TILE_PASS=1) or fixed-size tiles (default); empty rows,
    # long runs of tiny rows (more than 64 rows per work budget), rows of every light bin and heavier rows in between
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    monkeypatch.setenv("SPADA_B200_TILE_PASS", tile_pass)
    rng = np.random.default_rng(96)
    lens = rng.choice([0, 1, 2, 5, 11, 23, 40, 90, 300], size=6000, p=[.15, .25, .2, .15, .1, .07, .05, .02, .01])
    lens[1000:1400] = 1      # a stretch of tiny rows: the row cap of a tile, not the work budget, cuts here
    lens[2000:2100] = 0
    a = random_csr(6000, 3000, row_nnz=lens, seed=97, values="signed")
    for n_cols in (5000, 1 << 24):     # 32-bit and 64-bit keys
        b = random_csr(3000, n_cols, row_nnz=rng.integers(0, 24, size=3000), seed=98, values="signed")
        e = spada.Engine(single_pass=True)
        try:
            r = e.spgemm(a, b)
            st = r.stats()
            names = [L["name"] for L in st["launches"]]
            assert any(n.startswith("tile_pass" if tile_pass == "1" else "fused") for n in names), names
            check(r, oracle.spgemm(a, b, threads=oracle.max_threads()), True)
            if tile_pass == "1":
                assert st["bins"]["32"]["window"] == [48, 32] and st["bins"]["512"]["window"] == [3, 32]
        finally:
            e.close()


def test_narrow_outputs_both_long_row_paths(spada, oracle, cari, monkeypatch):
    # cari (400 output columns): the dense-accumulator path (default for B.cols <= 16384) and, with SPADA_B200_NO_DENSE=1,
    # the chunk-sort + merge path (every chunk carries every column: runs of ~361 equal columns across 36 chunks)
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    g = spada.GEMM.from_mat("cari", cari)
    ref = oracle.spgemm(g.a, g.b, threads=oracle.max_threads())
    for no_dense, kernel in (("", "long_dense"), ("1", "long_sort")):
        if no_dense:
            monkeypatch.setenv("SPADA_B200_NO_DENSE", no_dense)
        e = spada.Engine()
        try:
            r = e.spgemm(g.a, g.b)
            assert kernel in [L["name"] for L in r.stats()["launches"]]
            check(r, ref, True)
        finally:
            e.close()


def test_host_to_host_one_call(engine, oracle, spada):
    # host CSR in, whole C in host arrays, uploads / kernels / downloads overlapped: A != B (A goes up in row panels),
    # A is B (alias), an empty A, and a non-canonical A (reported once all of it is on the device)
    rng = np.random.default_rng(99)
    lens = rng.choice([0, 2, 9, 40, 150, 900], size=5000, p=[.1, .3, .3, .2, .07, .03])
    a = random_csr(5000, 1500, row_nnz=lens, seed=100, values="signed")
    b = random_csr(1500, 20000, row_nnz=rng.integers(0, 50, size=1500), seed=101, values="signed")
    sq = random_csr(1500, 1500, row_nnz=rng.integers(0, 30, size=1500), seed=102, values="signed")
    for x, y in ((a, b), (sq, sq), (sp.csr_matrix((7, 1500)), b)):
        ref = oracle.spgemm(x, y, threads=oracle.max_threads())
        ip = np.zeros(x.shape[0] + 1, dtype=np.int64); ix = np.zeros(len(ref[1]) + 3, dtype=np.int32); dx = np.zeros(len(ref[1]) + 3)
        st = engine.spgemm_host_to_host(x, y, ip, ix, dx)
        n = st["nnz_c"]
        assert n == len(ref[1]) and st["products"] == int(oracle.flops(x, y).sum())
        assert np.array_equal(ip, ref[0]) and np.array_equal(ix[:n], ref[1])
        assert np.array_equal(dx[:n].view(np.uint64), ref[2].view(np.uint64))
    bad = a.copy()
    rows2 = np.flatnonzero(np.diff(bad.indptr) >= 2)
    r = rows2[len(rows2) // 2]
    s0 = bad.indptr[r]
    bad.indices[s0], bad.indices[s0 + 1] = bad.indices[s0 + 1], bad.indices[s0]
    cap = int(oracle.flops(a, b).sum())
    ip = np.zeros(a.shape[0] + 1, dtype=np.int64); ix = np.zeros(cap, dtype=np.int32); dx = np.zeros(cap)
    with pytest.raises(spada.SpadaB200Error) as e:
        engine.spgemm_host_to_host(bad, b, ip, ix, dx)
    assert e.value.status == "UNSORTED_INPUT"
