"""Small shapes through every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from conftest import random_csr
import oracle
pkg = importlib.import_module("spada-sim_b200")
rng = np.random.default_rng(7)
for mode in ({"two_phase": True}, {"single_pass": True}):
    e = pkg.Engine(**mode)
    for ka, lb, n in [(3, 4, 300), (5, 5, 5000), (12, 2, 64), (8, 8, 5000), (16, 16, 5000), (20, 25, 5000), (32, 32, 5000),
                      (64, 64, 5000), (70, 100, 5000), (300, 230, 3000), (120, 300, 40),
                      # one bit short of 32-bit keys (top-bit split) on the 2048 / 4096 sorts and the chunks of long rows; 64-bit keys
                      (40, 50, 1 << 22), (64, 64, 1 << 21), (70, 100, 1 << 21), (64, 64, 1 << 23), (70, 100, 1 << 23)]:
        a = random_csr(40, 600, row_nnz=rng.integers(max(ka - 2, 0), ka + 1, size=40), seed=ka, values="signed")
        b = random_csr(600, n, row_nnz=np.minimum(rng.integers(max(lb - 2, 0), lb + 1, size=600), n), seed=lb, values="signed")
        r = e.spgemm(a, b)
        ip, ix, dx = r.to_host()
        ref = oracle.spgemm(a, b, threads=2)
        ok = np.array_equal(ip, ref[0]) and np.array_equal(ix, ref[1]) and np.array_equal(dx.view(np.uint64), ref[2].view(np.uint64))
        print(mode, ka, lb, n, "ok" if ok else "MISMATCH", flush=True)
    a = random_csr(300, 900, row_nnz=rng.integers(0, 30, size=300), seed=99)
    t = e.transpose(e.upload(a)).to_scipy()
    ref = a.T.tocsr(); ref.sort_indices()
    print("transpose", "ok" if (np.array_equal(t.indptr, ref.indptr) and np.array_equal(t.indices, ref.indices)
                                and np.array_equal(t.data, ref.data)) else "MISMATCH", flush=True)
    e.close()
