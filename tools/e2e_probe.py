#!/usr/bin/env python
"""Where the end-to-end time goes: upload, one-shot product + whole-C copy, row panels of several sizes."""
import ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("spada-sim_b200")
abi, lib = pkg._abi, pkg._abi.lib()
name = sys.argv[1] if len(sys.argv) > 1 else "rect"
a, b = bench.load_workload(pkg, name, 1.0)
eng = pkg.Engine(device=0)
def pin(arr):
    v, p = bench.pinned_copy(pkg, arr); return v
ai, ax, ad = pin(a.indptr.astype(np.int32)), pin(a.indices.astype(np.int32)), pin(a.data)
bi, bx, bd = (ai, ax, ad) if b is a else (pin(b.indptr.astype(np.int32)), pin(b.indices.astype(np.int32)), pin(b.data))
def view(m, i, x, d):
    return abi.CsrView32(m.shape[0], m.shape[1], m.nnz, i.ctypes.data_as(C.POINTER(C.c_int32)), x.ctypes.data_as(C.POINTER(C.c_int32)), d.ctypes.data_as(C.POINTER(C.c_double)))
va, vb = view(a, ai, ax, ad), view(b, bi, bx, bd)
def upload():
    pa, pb = C.c_void_p(), C.c_void_p()
    abi.check(lib.spada_b200_upload32(eng._h, C.byref(va), C.byref(pa)))
    abi.check(lib.spada_b200_upload32(eng._h, C.byref(vb), C.byref(pb)))
    return pa, pb
for _ in range(2):
    t = time.perf_counter(); pa, pb = upload(); t_up = time.perf_counter() - t
    lib.spada_b200_csr_free(pa); lib.spada_b200_csr_free(pb)
print(f"upload A+B: {t_up * 1e3:.1f} ms")
da, db = eng.upload(a), eng.upload(b)
r = eng.spgemm_dev(da, db); nnz = r.nnz; r.free()
op, oc, ov = pin(np.zeros(a.shape[0] + 1, dtype=np.int64)), pin(np.zeros(nnz, dtype=np.int32)), pin(np.zeros(nnz))
for it in range(3):
    t = time.perf_counter(); r = eng.spgemm_dev(da, db); t1 = time.perf_counter(); r.to_host(out=(op, oc, ov)); t2 = time.perf_counter(); r.free()
print(f"one shot: compute {1e3 * (t1 - t):.1f} ms + whole-C copy {1e3 * (t2 - t1):.1f} ms ({12 * nnz / (t2 - t1) / 1e9:.1f} GB/s)")
for pp in (1 << 40, 0, 1 << 24, 1 << 23):
    for it in range(2):
        t = time.perf_counter(); st = eng.spgemm_to_host(da, db, op, oc, ov, pp); dt = time.perf_counter() - t
    print(f"to_host panel_products={pp}: {st['panels']} panels, {dt * 1e3:.1f} ms ({12 * nnz / dt / 1e9:.1f} GB/s)")
