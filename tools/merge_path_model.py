#!/usr/bin/env python
"""Index-level model of the merge-based long-row path planned for the next round.

Not product code and not part of the test suite: it pins down, on the CPU, the two pieces of index logic the CUDA
kernels will need, so that GPU time next round goes into performance rather than off-by-one hunting.

1. ``cut_runs``: the A entries of one row are cut greedily into groups of at most ``cap`` intermediate products; a B
   row longer than ``cap`` is cut into pieces of ``cap`` elements.  Every group becomes one sorted, reduced partial
   row (on the GPU: k_bitonic_numeric_cta); at most 2 * ceil(p / cap) + 2 runs per row.
2. ``merge_tiled``: two sorted runs with unique columns are merged by independent tiles of ``tile`` consumed
   elements (merge path, ties X-before-Y).  Equal columns are summed, X's value first (= ascending k when runs are
   merged in order, the oracle's association).  An equal pair can straddle a tile boundary; both tiles decide
   that from the inputs alone: the tile that ends on the X element adds the next Y value, the next tile skips its
   first Y element.  Every tile reports how many outputs it wrote, so a scan places the tiles.

``python tools/merge_path_model.py`` runs randomized checks of both against dictionary arithmetic.
"""
import numpy as np


def cut_runs(b_lens, cap):
    """b_lens: B-row length of every A entry of the row.  Returns [(first_entry, last_entry_exclusive, off0, off1)]:
    whole entries [first, last) or, for a B row longer than cap, one entry with the element range [off0, off1)."""
    runs, start, acc = [], 0, 0
    for e, n in enumerate(b_lens):
        n = int(n)
        if n > cap:
            if acc:
                runs.append((start, e, 0, 0))
            for o in range(0, n, cap):
                runs.append((e, e + 1, o, min(n, o + cap)))
            start, acc = e + 1, 0
        elif acc + n > cap:
            runs.append((start, e, 0, 0))
            start, acc = e, n
        else:
            acc += n
    if acc:
        runs.append((start, len(b_lens), 0, 0))
    return runs


def merge_path(xc, yc, d):
    """Largest i with i <= d such that X[:i] and Y[:d-i] are the first d elements of the merge (ties X first)."""
    lo, hi = max(0, d - len(yc)), min(d, len(xc))
    while lo < hi:
        i = (lo + hi + 1) // 2
        # X[i-1] may be taken before Y[d-i] iff X[i-1] <= Y[d-i] (X first on ties)
        if d - i < len(yc) and xc[i - 1] > yc[d - i]:
            hi = i - 1
        else:
            lo = i
    return lo


def merge_tile(xc, xv, yc, yv, d0, d1):
    """One tile: consumes the merged elements [d0, d1).  Returns (cols, vals) it writes."""
    i, i1 = merge_path(xc, yc, d0), merge_path(xc, yc, d1)
    j, j1 = d0 - i, d1 - i1
    out_c, out_v = [], []
    # the first Y element belongs to the previous tile if it pairs with that tile's last X element
    if j < j1 and i > 0 and xc[i - 1] == yc[j]:
        j += 1
    while i < i1 or j < j1:
        take_x = j >= j1 or (i < i1 and xc[i] <= yc[j])
        if take_x:
            c, v = xc[i], xv[i]
            i += 1
            if j < j1 and yc[j] == c:          # pair inside the tile
                v = v + yv[j]
                j += 1
            elif j == j1 and j1 < len(yc) and i == i1 and yc[j1] == c:   # pair straddling the boundary: peek ahead
                v = v + yv[j1]
            out_c.append(c); out_v.append(v)
        else:
            out_c.append(yc[j]); out_v.append(yv[j])
            j += 1
    return out_c, out_v


def merge_tiled(xc, xv, yc, yv, tile):
    total = len(xc) + len(yc)
    cols, vals, counts = [], [], []
    for d0 in range(0, total, tile):
        c, v = merge_tile(xc, xv, yc, yv, d0, min(total, d0 + tile))
        counts.append(len(c)); cols += c; vals += v
    return np.array(cols, dtype=np.int64), np.array(vals), counts


def _check(rng):
    # runs
    for _ in range(200):
        cap = int(rng.integers(4, 64))
        lens = rng.integers(0, 3 * cap, size=int(rng.integers(1, 60)))
        runs = cut_runs(lens, cap)
        covered = np.zeros(int(lens.sum()), dtype=np.int64)
        starts = np.concatenate([[0], np.cumsum(lens)])
        for f, l, o0, o1 in runs:
            if o1:
                assert l == f + 1 and 0 < o1 - o0 <= cap
                covered[starts[f] + o0:starts[f] + o1] += 1
            else:
                assert 0 < lens[f:l].sum() <= cap
                covered[starts[f]:starts[l]] += 1
        assert (covered == 1).all()
        assert len(runs) <= 2 * -(-int(lens.sum()) // cap) + 2
    # tiled merge
    for _ in range(400):
        n = int(rng.integers(1, 60))
        xc = np.sort(rng.choice(n * 2, size=int(rng.integers(0, n)), replace=False))
        yc = np.sort(rng.choice(n * 2, size=int(rng.integers(0, n)), replace=False))
        xv, yv = rng.uniform(-1, 1, len(xc)), rng.uniform(-1, 1, len(yc))
        ref = {}
        for c, v in zip(xc, xv):
            ref[int(c)] = v
        for c, v in zip(yc, yv):
            ref[int(c)] = ref[int(c)] + v if int(c) in ref else v
        for tile in (1, 2, 3, 7, 16, 1000):
            c, v, counts = merge_tiled(xc, xv, yc, yv, tile)
            assert list(c) == sorted(ref), (xc, yc, tile, c)
            assert all(v[k] == ref[int(c[k])] for k in range(len(c)))      # bit-identical: X's value first
            assert sum(counts) == len(c)


if __name__ == "__main__":
    _check(np.random.default_rng(0))
    print("merge-path model: run cutting and tiled merge agree with dictionary arithmetic")
