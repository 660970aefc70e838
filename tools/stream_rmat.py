#!/usr/bin/env python
"""C larger than HBM: Graph500-skew R-MAT (0.57, 0.19, 0.19, 0.05) squared through the row-panel stream.

SURVEY.md 8d measured this operand at 55.4 G intermediate products and an estimated 26 G nnz(C) (~310 GB) for scale 21 --
more than one B200 holds.  `spada_b200_spgemm_stream` computes C in row panels, copies every finished panel to pinned host
memory while the next one is computed and hands it to a sink; the sink here keeps a checksum (nnz, sum of column ids, sum
of values, a digest of the row pointers) and compares a sample of rows with the CPU oracle, then drops the panel.

    python tools/stream_rmat.py [scale] [panel_products]
"""
import hashlib, importlib, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402  (the checker of the sampled rows only)

pkg = importlib.import_module("spada-sim_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 21
panel = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 28

t = time.time()
a = pkg.workloads.rmat(scale, 16, seed=42, probs=(0.57, 0.19, 0.19, 0.05))
print(f"operand: R-MAT scale {scale}, Graph500 skew, {a.shape[0]} rows, nnz {a.nnz} ({time.time() - t:.0f} s to generate)", flush=True)
f = oracle.flops(a, a)
total = int(f.sum())
print(f"intermediate products {total} (upper bound of nnz(C): {12 * total / 1e9:.0f} GB of C), heaviest row {int(f.max())}", flush=True)

rng = np.random.default_rng(3)
sample = np.sort(rng.choice(a.shape[0], 3000, replace=False))
sample = sample[f[sample] < 5_000_000]           # keep the oracle's part short
ref_p, ref_c, ref_v = oracle.spgemm(a[sample], a, threads=oracle.max_threads())
state = {"nnz": 0, "col_sum": 0, "val_sum": 0.0, "panels": 0, "checked": 0, "digest": hashlib.sha256(), "bad": 0}


def sink(r0, r1, n0, ip, ix, dx):
    state["nnz"] += len(ix)
    state["col_sum"] += int(ix.sum(dtype=np.int64))
    state["val_sum"] += float(dx.sum())
    state["digest"].update(np.diff(ip).astype(np.int32).tobytes())
    state["panels"] += 1
    lo, hi = np.searchsorted(sample, [r0, r1])
    for k in range(lo, hi):                       # sampled rows of this panel against the oracle, bit for bit
        r = sample[k] - r0
        s, e = ip[r] - n0, ip[r + 1] - n0
        cs, ce = ref_p[k], ref_p[k + 1]
        ok = (e - s == ce - cs) and np.array_equal(ix[s:e], ref_c[cs:ce]) and np.array_equal(dx[s:e].view(np.uint64), ref_v[cs:ce].view(np.uint64))
        state["checked"] += 1
        state["bad"] += 0 if ok else 1


eng = pkg.Engine(device=0)
da = eng.upload(a)
t = time.time()
st = eng.spgemm_stream(da, da, sink, panel)
dt = time.time() - t
print(f"streamed C: {st['panels']} panels, nnz(C) {st['nnz_c']} = {12 * st['nnz_c'] / 1e9:.1f} GB (device memory: 180 GB), {dt:.1f} s wall, "
      f"{2 * total / dt / 1e9:.1f} GFLOP/s end to end incl. the host-side checksum", flush=True)
print(f"checksum: nnz {state['nnz']} col_sum {state['col_sum']} val_sum {state['val_sum']!r} row_len_sha256 {state['digest'].hexdigest()[:16]}")
print(f"sampled rows against the oracle: {state['checked']} rows, {state['bad']} differ")
assert state["nnz"] == st["nnz_c"] and state["bad"] == 0 and state["checked"] == len(sample)
