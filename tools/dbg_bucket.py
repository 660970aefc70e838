import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_csr
pkg = importlib.import_module("spada-sim_b200")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle
e = pkg.Engine(two_phase=True)
for ka, lb in [(32, 32), (40, 50), (64, 64), (70, 100), (300, 40)]:
    a = random_csr(60, 600, row_nnz=ka, seed=ka * 7 + lb)
    b = random_csr(600, 5000, row_nnz=lb, seed=ka * 11 + lb)
    r = e.spgemm(a, b)
    ip, ix, dx = r.to_host()
    ref = oracle.spgemm(a, b, threads=4)
    print(ka, lb, "ptr", np.array_equal(ip, ref[0]), "col", np.array_equal(ix, ref[1]) if len(ix) == len(ref[1]) else False,
          "maxrel", float(np.max(np.abs(dx - ref[2]) / np.abs(ref[2]))) if len(dx) == len(ref[2]) else None, flush=True)
