"""Python host mirror over the C ABI: handles, device operands, results.

Everything here is plumbing around libspada_b200.so (ctypes); all arithmetic happens in the
CUDA kernels.  The object model follows include/spada_b200.h one to one.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import scipy.sparse as sp

from . import _abi
from ._abi import check, lib

U64_MAX = (1 << 64) - 1
# bin b holds rows with at most 32 * 2^(b-1) intermediate products; bins above 4096 are the LONG rows (chunk sorts + merge levels)
BIN_NAMES = ["empty"] + [str(32 << (b - 1)) for b in range(1, 29)]
LONG_BINS = BIN_NAMES[9:]


def _ptr(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class DeviceCsr:
    """A CSR operand resident on the device (spada_b200_csr_t)."""

    def __init__(self, engine: "Engine", handle: C.c_void_p, keepalive=None):
        self.engine = engine
        self._h = handle
        self._keep = keepalive
        r, c, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().spada_b200_csr_shape(self._h, C.byref(r), C.byref(c), C.byref(n)))
        self.shape = (r.value, c.value)
        self.nnz = n.value

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().spada_b200_csr_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def to_scipy(self) -> sp.csr_matrix:
        """Copies the operand back to the host (int64 indptr, int32 indices, float64 data)."""
        ip = np.empty(self.shape[0] + 1, dtype=np.int64)
        ix = np.empty(max(self.nnz, 1), dtype=np.int32)
        dx = np.empty(max(self.nnz, 1), dtype=np.float64)
        check(lib().spada_b200_csr_download32(self._h, _ptr(ip, C.c_int64), _ptr(ix, C.c_int32), _ptr(dx, C.c_double)))
        return sp.csr_matrix((dx[:self.nnz], ix[:self.nnz], ip), shape=self.shape)

    def set_one_shot(self):
        """No fiber store for this operand (it is used for one product only)."""
        check(lib().spada_b200_csr_set_one_shot(self._h))

    def prepare(self) -> float:
        """Builds the fiber store the kernels gather B rows from (automatic for uploaded operands on their first
        use as B; needed for wrapped device arrays).  Returns the device time in ms."""
        ms = C.c_float(0.0)
        check(lib().spada_b200_csr_prepare(self.engine._h, self._h, C.byref(ms)))
        return ms.value

    def free(self):
        if self._h is not None:
            if self.engine._h is not None:   # a destroyed engine already released every device block
                lib().spada_b200_csr_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Result:
    """C = A x B resident on the device (spada_b200_result_t)."""

    def __init__(self, engine: "Engine", handle: C.c_void_p):
        self.engine = engine
        self._h = handle
        r, c, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().spada_b200_result_shape(self._h, C.byref(r), C.byref(c), C.byref(n)))
        self.shape = (r.value, c.value)
        self.nnz = n.value

    def stats(self) -> dict:
        st = _abi.Stats()
        check(lib().spada_b200_result_stats(self._h, C.byref(st)))
        return _stats_dict(st)

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().spada_b200_result_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def to_host(self, out=None):
        """(indptr int64, indices int32, data float64) on the host; `out` = preallocated triple."""
        if out is None:
            out = (np.empty(self.shape[0] + 1, dtype=np.int64), np.empty(self.nnz, dtype=np.int32),
                   np.empty(self.nnz, dtype=np.float64))
        ip, ix, dx = out
        check(lib().spada_b200_result_copy32(self._h, _ptr(ip, C.c_int64), _ptr(ix, C.c_int32), _ptr(dx, C.c_double)))
        return ip, ix, dx

    def to_host_usize(self):
        """The reference's layout: usize indptr / usize indices / f64 data (storage.rs:150-160)."""
        ip = np.empty(self.shape[0] + 1, dtype=np.uint64)
        ix = np.empty(self.nnz, dtype=np.uint64)
        dx = np.empty(self.nnz, dtype=np.float64)
        check(lib().spada_b200_result_copy(self._h, _ptr(ip, C.c_uint64), _ptr(ix, C.c_uint64), _ptr(dx, C.c_double)))
        return ip, ix, dx

    def to_scipy(self) -> sp.csr_matrix:
        ip, ix, dx = self.to_host()
        return sp.csr_matrix((dx, ix, ip), shape=self.shape)

    def free(self):
        if self._h is not None:
            if self.engine._h is not None:
                lib().spada_b200_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _stats_dict(st) -> dict:
    out = {k: getattr(st, k) for k in ("rows", "cols", "nnz_a", "nnz_b", "products", "nnz_c", "ms_total",
                                        "ms_flops", "ms_symbolic", "ms_scan", "ms_numeric", "ms_h2d", "ms_d2h",
                                        "n_launches")}
    out["bins"] = {BIN_NAMES[i]: {"rows": st.bin_rows[i], "products": st.bin_products[i],
                                  "window": [st.bin_window_rows[i], st.bin_window_lanes[i]]}
                   for i in range(len(BIN_NAMES)) if st.bin_rows[i]}
    out["launches"] = [{"name": st.launches[i].name.decode(), "ms": st.launches[i].ms,
                        "grid": st.launches[i].grid, "rows": st.launches[i].rows,
                        "products": st.launches[i].products} for i in range(st.n_recorded)]
    return out


class CBuf:
    """Full-size C buffers on one GPU (spada_b200_cbuf_t): the destination of sharded products."""

    def __init__(self, engine: "Engine", handle, rows, cols, capacity, imported=False):
        self.engine, self._h = engine, handle
        self.shape, self.capacity, self.imported = (rows, cols), capacity, imported

    def export(self) -> bytes:
        buf = C.create_string_buffer(3 * _abi.IPC_HANDLE_BYTES)
        check(lib().spada_b200_cbuf_export(self._h, buf))
        return buf.raw

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().spada_b200_cbuf_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @property
    def nnz(self) -> int:
        n = C.c_uint64()
        check(lib().spada_b200_cbuf_nnz(self._h, C.byref(n)))
        return n.value

    def to_host(self):
        n = self.nnz
        ip = np.empty(self.shape[0] + 1, dtype=np.int64)
        ix = np.empty(n, dtype=np.int32)
        dx = np.empty(n, dtype=np.float64)
        check(lib().spada_b200_cbuf_copy32(self._h, _ptr(ip, C.c_int64), _ptr(ix, C.c_int32), _ptr(dx, C.c_double)))
        return ip, ix, dx

    def free(self):
        if self._h is not None:
            if self.engine._h is not None:
                lib().spada_b200_cbuf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Shard:
    """A row shard's product between its two halves (spada_b200_shard_t)."""

    def __init__(self, engine: "Engine", handle, nnz_local):
        self.engine, self._h, self.nnz_local = engine, handle, nnz_local

    def finish(self, bufs, nnz_offset: int = 0, d_shard_nnz: int = 0, shard_index: int = 0) -> dict:
        """Stores the shard's rows into every buffer of ``bufs`` (own first, then the peers') at its global offset."""
        arr = (C.c_void_p * len(bufs))(*[b._h for b in bufs])
        st = _abi.Stats()
        h, self._h = self._h, None
        check(lib().spada_b200_shard_finish(h, arr, len(bufs), nnz_offset, d_shard_nnz or None, shard_index, C.byref(st)))
        return _stats_dict(st)

    def __del__(self):
        try:
            if self._h is not None and self.engine._h is not None:
                lib().spada_b200_shard_abort(self._h)
            self._h = None
        except Exception:
            pass


class Group:
    """All GPUs of this process behind one call (spada_b200_group_t): what the single-process CLI drives."""

    def __init__(self, n_gpus: int, accelerator: str = "spada", lane_num: int = 8, block_shape=(1, 10000000),
                 validate: bool = True):
        opts = _abi.Opts()
        opts.device = 0
        opts.accelerator = _abi.ACCELERATORS[accelerator.lower()]
        opts.lane_num = lane_num
        opts.block_shape[0] = min(int(block_shape[0]), 0xffffffff)
        opts.block_shape[1] = min(int(block_shape[1]), 0xffffffff)
        opts.flags = _abi.FLAG_VALIDATE if validate else 0
        h = C.c_void_p()
        check(lib().spada_b200_group_create(C.byref(opts), n_gpus, C.byref(h)))
        self._h = h
        self.n_gpus = n_gpus

    def spgemm(self, a: sp.csr_matrix, b: sp.csr_matrix) -> "Result":
        out = C.c_void_p()
        va, ka = Engine._view32(a)
        vb, kb = (va, ka) if b is a else Engine._view32(b)
        check(lib().spada_b200_group_spgemm32(self._h, C.byref(va), C.byref(vb), C.byref(out)))
        return Result(self, out)

    def close(self):
        if self._h is not None:
            lib().spada_b200_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One engine handle = one device + stream + memory pool (spada_b200_t)."""

    def __init__(self, device: int = -1, accelerator: str = "spada", lane_num: int = 8,
                 block_shape=(1, 10000000), validate: bool = True, stream: Optional[int] = None,
                 two_phase: bool = False, single_pass: bool = False, serial: bool = False):
        opts = _abi.Opts()
        opts.device = device
        opts.accelerator = _abi.ACCELERATORS[accelerator.lower()]
        opts.lane_num = lane_num
        opts.block_shape[0] = min(int(block_shape[0]), 0xffffffff)
        opts.block_shape[1] = min(int(block_shape[1]), 0xffffffff)
        opts.flags = (_abi.FLAG_VALIDATE if validate else 0) | (_abi.FLAG_TWO_PHASE if two_phase else 0) | \
            (_abi.FLAG_SINGLE_PASS if single_pass else 0) | (_abi.FLAG_SERIAL if serial else 0)
        opts.stream = stream
        h = C.c_void_p()
        check(lib().spada_b200_create(C.byref(opts), C.byref(h)))
        self._h = h
        # the stream the kernels run on, None = engine-owned (also for handle 0, CUDA's legacy default stream -- what
        # torch.cuda.current_stream().cuda_stream is outside a stream context: the ABI reads NULL as "engine-owned")
        self.stream = stream if stream else None

    # -- lifetime --
    def close(self):
        if self._h is not None:
            lib().spada_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream: Optional[int]):
        check(lib().spada_b200_set_stream(self._h, stream))
        self.stream = stream if stream else None

    def synchronize(self):
        check(lib().spada_b200_synchronize(self._h))

    def trim(self):
        check(lib().spada_b200_trim(self._h))

    # -- operands --
    @staticmethod
    def _view32(m: sp.csr_matrix):
        ip = np.ascontiguousarray(m.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(m.indices, dtype=np.int32)
        dx = np.ascontiguousarray(m.data, dtype=np.float64)
        v = _abi.CsrView32(m.shape[0], m.shape[1], int(ip[-1]) if len(ip) else 0, _ptr(ip, C.c_int32),
                           _ptr(ix, C.c_int32), _ptr(dx, C.c_double))
        return v, (ip, ix, dx)

    @staticmethod
    def _view64(m: sp.csr_matrix):
        ip = np.ascontiguousarray(m.indptr, dtype=np.uint64)
        ix = np.ascontiguousarray(m.indices, dtype=np.uint64)
        dx = np.ascontiguousarray(m.data, dtype=np.float64)
        v = _abi.CsrView(m.shape[0], m.shape[1], int(ip[-1]) if len(ip) else 0, _ptr(ip, C.c_uint64),
                         _ptr(ix, C.c_uint64), _ptr(dx, C.c_double))
        return v, (ip, ix, dx)

    def upload(self, m: sp.csr_matrix, usize: bool = False) -> DeviceCsr:
        """Host CSR -> device.  usize=True goes through the reference's Vec<usize> layout."""
        out = C.c_void_p()
        if usize or m.nnz >= (1 << 31):
            v, keep = self._view64(m)
            check(lib().spada_b200_upload(self._h, C.byref(v), C.byref(out)))
        else:
            v, keep = self._view32(m)
            check(lib().spada_b200_upload32(self._h, C.byref(v), C.byref(out)))
        return DeviceCsr(self, out)

    def transpose(self, a: DeviceCsr) -> DeviceCsr:
        """B = A^T on the device (GEMM::from_mat's transpose for non-square SS workloads, gemm.rs:41-53)."""
        out = C.c_void_p()
        check(lib().spada_b200_transpose(self._h, a._h, C.byref(out)))
        return DeviceCsr(self, out)

    def wrap_device(self, rows, cols, nnz, d_indptr: int, d_indices: int, d_data: int, keepalive=None) -> DeviceCsr:
        out = C.c_void_p()
        check(lib().spada_b200_csr_wrap_device(self._h, rows, cols, nnz, d_indptr, d_indices, d_data, C.byref(out)))
        return DeviceCsr(self, out, keepalive)

    # -- the hot path --
    def spgemm_dev(self, a: DeviceCsr, b: DeviceCsr, row_begin: int = 0, row_end: Optional[int] = None) -> Result:
        out = C.c_void_p()
        check(lib().spada_b200_spgemm_dev(self._h, a._h, b._h, row_begin, U64_MAX if row_end is None else row_end,
                                          C.byref(out)))
        return Result(self, out)

    def spgemm(self, a: sp.csr_matrix, b: sp.csr_matrix, usize: bool = False) -> Result:
        """Host operands in, device result out, through the host-level ABI entry point."""
        out = C.c_void_p()
        if usize:
            va, ka = self._view64(a)
            vb, kb = (va, ka) if b is a else self._view64(b)
            check(lib().spada_b200_spgemm(self._h, C.byref(va), C.byref(vb), C.byref(out)))
        else:
            va, ka = self._view32(a)
            vb, kb = (va, ka) if b is a else self._view32(b)
            check(lib().spada_b200_spgemm32(self._h, C.byref(va), C.byref(vb), C.byref(out)))
        return Result(self, out)

    # -- row-panel streaming --
    def spgemm_stream(self, a: DeviceCsr, b: DeviceCsr, sink, panel_products: int = 0) -> dict:
        """C in row panels; ``sink(row_begin, row_end, nnz_begin, indptr, indices, data)`` gets numpy views that are valid
        only during the call (indptr holds global offsets).  The D2H copy of a panel overlaps the next panel's kernels."""
        err = []

        def tramp(_user, r0, r1, n0, ip, ix, dx):
            try:
                rows = r1 - r0
                p = np.ctypeslib.as_array(ip, shape=(rows + 1,))
                nnz = int(p[-1] - p[0])
                c = np.ctypeslib.as_array(ix, shape=(max(nnz, 1),))[:nnz]
                v = np.ctypeslib.as_array(dx, shape=(max(nnz, 1),))[:nnz]
                sink(int(r0), int(r1), int(n0), p, c, v)
                return 0
            except Exception as e:   # never unwind through the C frames
                err.append(e)
                return 1
        cb = _abi.PANEL_SINK(tramp)
        st = _abi.StreamStats()
        rc = lib().spada_b200_spgemm_stream(self._h, a._h, b._h, panel_products, cb, None, C.byref(st))
        if err:
            raise err[0]
        check(rc)
        return {k: getattr(st, k) for k in ("panels", "products", "nnz_c", "max_panel_products", "ms_total")}

    def spgemm_to_host(self, a: DeviceCsr, b: DeviceCsr, indptr: np.ndarray, indices: np.ndarray, data: np.ndarray,
                       panel_products: int = 0) -> dict:
        """The whole C into caller-allocated host arrays (int64 / int32 / float64; pinned for full PCIe speed)."""
        assert indptr.dtype == np.int64 and indices.dtype == np.int32 and data.dtype == np.float64
        assert len(indptr) >= a.shape[0] + 1 and len(indices) == len(data)
        st = _abi.StreamStats()
        check(lib().spada_b200_spgemm_to_host(self._h, a._h, b._h, panel_products, _ptr(indptr, C.c_int64),
                                              _ptr(indices, C.c_int32), _ptr(data, C.c_double), len(indices), C.byref(st)))
        return {k: getattr(st, k) for k in ("panels", "products", "nnz_c", "max_panel_products", "ms_total")}

    def spgemm_host_to_host(self, a: sp.csr_matrix, b: sp.csr_matrix, indptr: np.ndarray, indices: np.ndarray,
                            data: np.ndarray) -> dict:
        """Host CSR operands in, the whole C in host arrays, uploads / kernels / downloads overlapped (one ABI call)."""
        va, ka = self._view32(a)
        vb, kb = (va, ka) if b is a else self._view32(b)
        st = _abi.StreamStats()
        check(lib().spada_b200_spgemm32_host_to_host(self._h, C.byref(va), C.byref(vb), _ptr(indptr, C.c_int64),
                                                     _ptr(indices, C.c_int32), _ptr(data, C.c_double), len(indices), C.byref(st)))
        return {k: getattr(st, k) for k in ("panels", "products", "nnz_c", "max_panel_products", "ms_total")}

    # -- sharded runs --
    def cbuf_create(self, rows: int, cols: int, capacity_nnz: int) -> CBuf:
        out = C.c_void_p()
        check(lib().spada_b200_cbuf_create(self._h, rows, cols, capacity_nnz, C.byref(out)))
        return CBuf(self, out, rows, cols, capacity_nnz)

    def cbuf_import(self, handles: bytes, rows: int, cols: int, capacity_nnz: int) -> CBuf:
        out = C.c_void_p()
        buf = C.create_string_buffer(handles, len(handles))
        check(lib().spada_b200_cbuf_import(self._h, buf, rows, cols, capacity_nnz, C.byref(out)))
        return CBuf(self, out, rows, cols, capacity_nnz, imported=True)

    def shard_begin(self, a: DeviceCsr, b: DeviceCsr, row_begin: int, row_end: int, d_nnz_local: int = 0,
                    host_nnz: bool = False) -> Shard:
        """First half of a shard's product; the shard's nnz lands at device address ``d_nnz_local``."""
        out = C.c_void_p()
        n = C.c_uint64()
        check(lib().spada_b200_shard_begin(self._h, a._h, b._h, row_begin, row_end, d_nnz_local or None,
                                           C.byref(n) if host_nnz else None, C.byref(out)))
        return Shard(self, out, n.value if host_nnz else None)

    def flops(self, a: DeviceCsr, b: DeviceCsr, per_row: bool = False):
        total = C.c_uint64()
        arr = np.zeros(a.shape[0], dtype=np.uint64) if per_row else None
        check(lib().spada_b200_flops(self._h, a._h, b._h, C.byref(total), _ptr(arr, C.c_uint64) if per_row else None))
        return (total.value, arr) if per_row else total.value

    def plan_shards(self, a: DeviceCsr, b: DeviceCsr, n_shards: int) -> np.ndarray:
        bounds = np.zeros(n_shards + 1, dtype=np.uint64)
        check(lib().spada_b200_plan_shards(self._h, a._h, b._h, n_shards, _ptr(bounds, C.c_uint64)))
        return bounds.astype(np.int64)


def device_count() -> int:
    n = C.c_int()
    rc = lib().spada_b200_device_count(C.byref(n))
    return n.value if rc == 0 else 0
