"""Index-level models of device-side layout and ordering rules (numpy, no GPU).

The kernels themselves are checked against the oracle by the ``-m gpu`` tests; these models pin the *arguments* the
kernels rest on, so a change of a constant or a map shows up on the CPU first:

* ``KeySlot`` (csrc/cta_common.cuh): the XOR map of the CTA-wide sort's keys is a bijection, keeps linear accesses
  conflict-free and makes the 16-byte register<->shared-memory moves conflict-free;
* the padded staging of ``k_long_merge`` (csrc/longrow.cu: mg_col_at / mg_val_at);
* ``cta_split_top``: sorting (column without its top bit, arrival) and then splitting stably by the top bit is the
  (column, arrival) order -- the order the oracle sums in (oracle/spgemm_oracle.c);
* the last merge level's head count: tiles of 4096 outputs, the column before a tile taken as the larger of the two
  elements the merge path leaves behind (k_long_partition), add up to the row's nnz.
"""
import numpy as np
import pytest

CTA_THREADS = 256
BANKS = 32


def key_slot(e, key_bytes, n):
    le = 2 if key_bytes == 4 else 1
    u = (n // CTA_THREADS) * key_bytes // 16
    if u <= 2:                       # KeySlot::ON
        return e
    return e ^ (((e >> (le + 3)) & (u - 1)) << le)


@pytest.mark.parametrize("key_bytes,n", [(4, 1024), (4, 2048), (4, 4096), (8, 1024), (8, 2048), (8, 4096)])
def test_key_slot_map(key_bytes, n):
    e = np.arange(n)
    slot = np.array([key_slot(int(x), key_bytes, n) for x in e])
    assert np.array_equal(np.sort(slot), e)                        # a permutation of the key array
    per_vec = 16 // key_bytes
    assert np.array_equal(slot % per_vec, e % per_vec)             # whole 16-byte vectors move
    words = key_bytes // 4
    # consecutive keys read by the 32 lanes of a warp: every 4-byte word of a 128-byte transaction on its own bank
    for e0 in range(0, n, 32 // words):
        lanes = slot[e0:e0 + 32 // words]
        banks = np.concatenate([(lanes * words + w) % BANKS for w in range(words)])
        assert len(set(banks.tolist())) == len(banks)
    # a lane's E consecutive keys as 16-byte vectors: the 8 lanes of a quarter-warp hit 8 different bank groups
    ekeys = n // CTA_THREADS
    u = ekeys * key_bytes // 16
    if u > 2:
        for first_lane in range(0, CTA_THREADS, 8):
            for i in range(u):
                groups = {(slot[(first_lane + l) * ekeys + i * per_vec] // per_vec) % 8 for l in range(8)}
                assert len(groups) == 8


def test_merge_staging_padding():
    items, threads = 8, 512          # MG_ITEMS, MG_THREADS
    col_at = lambda e: e + (e >> 5)
    val_at = lambda e: e + (e >> 4)
    for warp in range(threads // 32):
        for q in range(items):
            e = np.array([(warp * 32 + lane) * items + q for lane in range(32)])
            assert len(set((col_at(e) % BANKS).tolist())) == 32                     # 32 lanes, 32 banks
            for half in (e[:16], e[16:]):                                           # 8-byte accesses go by half-warps
                assert len(set((val_at(half) % 16).tolist())) == 16
    e = np.arange(4096)
    assert col_at(e).max() < 4096 + 4096 // 32 + 16 and val_at(e).max() < 4096 + 4096 // 16 + 16   # MergeStage sizes
    assert len(set(col_at(e).tolist())) == 4096 and len(set(val_at(e).tolist())) == 4096


@pytest.mark.parametrize("n,cnt,seed", [(4096, 4096, 0), (4096, 3001, 1), (2048, 2048, 2), (1024, 700, 3), (4096, 1, 4)])
def test_split_by_top_bit_is_the_full_order(n, cnt, seed):
    rng = np.random.default_rng(seed)
    sb = n.bit_length() - 1
    width = 32 - sb                                  # column bits a 32-bit key holds next to the arrival index
    pool = np.concatenate([rng.integers(0, 40, 30), (1 << width) + rng.integers(0, 40, 30), [(1 << width) - 1, (1 << (width + 1)) - 1]])
    col = rng.choice(pool, cnt)                      # few distinct columns: long runs on both sides of the bit
    arrival = np.arange(cnt)
    key = ((col & ((1 << width) - 1)) << sb) | arrival            # what the expansion packs (the top bit falls off)
    assert key.max() < 1 << 32
    top = (col >> width) & 1                                      # what top_bit_mark records, by arrival
    order = np.argsort(key, kind="stable")                        # the network's result (keys are unique)
    t = arrival[order]
    low = top[t] == 0
    split = np.concatenate([t[low], t[~low]])                     # cta_split_top: stable, clear-bit products first
    want = np.lexsort((arrival, col))                             # (column, arrival): the oracle's summation order
    assert np.array_equal(split, want)
    n0 = int(low.sum())
    rebuilt = (key[split] >> sb) | ((np.arange(cnt) >= n0).astype(np.int64) << width)   # consumers: top bit iff i >= n0
    assert np.array_equal(rebuilt, col[split])


def merge_path(x, y, d):
    lo, hi = max(d - len(y), 0), min(d, len(x))
    while lo < hi:
        mid = (lo + hi) // 2
        if x[mid] <= y[d - 1 - mid]:
            lo = mid + 1
        else:
            hi = mid
    return lo


@pytest.mark.parametrize("nx,ny,ncols,seed", [(8192, 8192, 50, 0), (8192, 5000, 3000, 1), (4096, 1, 7, 2), (16384, 9000, 1, 3)])
def test_last_level_head_count(nx, ny, ncols, seed):
    unit = 4096
    rng = np.random.default_rng(seed)
    x = np.sort(rng.integers(0, ncols, nx))
    y = np.sort(rng.integers(0, ncols, ny))
    merged = np.concatenate([x, y])[np.argsort(np.concatenate([x, y]), kind="stable")]   # ties: X first
    total = 0
    for o0 in range(0, nx + ny, unit):
        o1 = min(o0 + unit, nx + ny)
        i0, i1 = merge_path(x, y, o0), (nx if o1 == nx + ny else merge_path(x, y, o1))
        j0, j1 = o0 - i0, o1 - i1
        tile = np.concatenate([x[i0:i1], y[j0:j1]])
        tile = tile[np.argsort(tile, kind="stable")]
        assert np.array_equal(tile, merged[o0:o1])
        prev = -1                                                    # k_long_partition, level == L
        if i0 > 0:
            prev = int(x[i0 - 1])
        if j0 > 0:
            prev = max(prev, int(y[j0 - 1]))
        assert prev == (int(merged[o0 - 1]) if o0 else -1)
        total += int((tile != np.concatenate([[prev], tile[:-1]])).sum())   # k_long_merge: heads among the outputs
    assert total == len(np.unique(merged))
