"""Regenerates tests/golden/* from the reference's only shipped operand.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py

Outputs
  cari_csr.npz              A = scipy.io.mmread('matrices/cari.mtx').tocsr()  (py2rust.rs:63-80)
                            stored as int32 indptr/indices + float64 data, exactly what scipy
                            hands to pyo3 in the reference.
  cari_known_answers.json   structure/values of C = A x A^T (gemm.rs:41-53: cari is 400x1200,
                            not square => B = A^T) computed with scipy `A @ B` +
                            sort_indices(), the independent cross-check of SURVEY.md 8c.
                            The sha256 prefixes / sum reproduce the survey's known answers.
"""
import hashlib
import json
import os

import numpy as np
import scipy.io as spio

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    with open(os.path.join(REF, "matrices", "cari.mtx"), "r") as f:
        a = spio.mmread(f).tocsr()
    a.sort_indices()
    np.savez_compressed(os.path.join(HERE, "cari_csr.npz"), shape=np.array(a.shape, dtype=np.int64),
                        indptr=a.indptr.astype(np.int32), indices=a.indices.astype(np.int32),
                        data=a.data.astype(np.float64))
    b = a.T.tocsr()
    b.sort_indices()
    c = (a @ b).tocsr()
    c.sort_indices()
    sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    known = {
        "source": "matrices/cari.mtx (reference), C = A x A^T via scipy %s" % __import__("scipy").__version__,
        "a_shape": list(a.shape), "a_nnz": int(a.nnz),
        "b_shape": list(b.shape), "b_nnz": int(b.nnz),
        "products": int(np.diff(b.indptr)[a.indices].sum()),
        "c_shape": list(c.shape), "c_nnz": int(c.nnz),
        "c_data_sum": float(c.data.sum()),
        "c_min_abs": float(np.abs(c.data).min()),
        "sha256_indptr_i64": sha(c.indptr.astype("<i8")),
        "sha256_indices_i64": sha(c.indices.astype("<i8")),
        "sha256_data_f64": sha(c.data.astype("<f8")),
        "row0_cols": c.indices[c.indptr[0]:c.indptr[0] + 5].tolist(),
        "row0_vals": c.data[c.indptr[0]:c.indptr[0] + 5].tolist(),
        "a_head": {"data": a.data[:5].tolist(), "indices": a.indices[:5].tolist(),
                   "indptr": a.indptr[:5].tolist()},
    }
    with open(os.path.join(HERE, "cari_known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)
    print(json.dumps(known, indent=1))


if __name__ == "__main__":
    main()
