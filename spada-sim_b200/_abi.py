"""ctypes binding of libspada_b200.so -- the C ABI of include/spada_b200.h.

The library is the product; this module only declares its entry points.  It fails loudly when
the shared object is missing (no CPU fallback anywhere in the package).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPADA_B200_LIB lets a tuning run load an alternative build of the same library (never a different backend)
LIB_PATH = os.environ.get("SPADA_B200_LIB") or os.path.join(_HERE, "lib", "libspada_b200.so")

MAX_BINS = 32
MAX_LAUNCHES = 64
ABI_VERSION = 2
IPC_HANDLE_BYTES = 64

STATUS = {0: "OK", 1: "INVALID_ARG", 2: "UNSORTED_INPUT", 3: "DIM_MISMATCH", 4: "CUDA_ERROR",
          5: "NCCL_ERROR", 6: "OOM", 7: "NO_DEVICE", 8: "TOO_LARGE"}
ACCELERATORS = {"ip": 0, "op": 1, "multirow": 2, "spada": 3}
FLAG_VALIDATE = 1
FLAG_TWO_PHASE = 2
FLAG_SINGLE_PASS = 4
FLAG_SERIAL = 8


class CsrView(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz", C.c_uint64),
                ("indptr", C.POINTER(C.c_uint64)), ("indices", C.POINTER(C.c_uint64)),
                ("data", C.POINTER(C.c_double))]


class CsrView32(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz", C.c_uint64),
                ("indptr", C.POINTER(C.c_int32)), ("indices", C.POINTER(C.c_int32)),
                ("data", C.POINTER(C.c_double))]


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("accelerator", C.c_int32), ("lane_num", C.c_uint32),
                ("block_shape", C.c_uint32 * 2), ("flags", C.c_uint32), ("stream", C.c_void_p)]


class Launch(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_float), ("grid", C.c_uint32),
                ("rows", C.c_uint64), ("products", C.c_uint64), ("nnz", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz_a", C.c_uint64), ("nnz_b", C.c_uint64),
                ("products", C.c_uint64), ("nnz_c", C.c_uint64),
                ("bin_rows", C.c_uint64 * MAX_BINS), ("bin_products", C.c_uint64 * MAX_BINS),
                ("bin_window_rows", C.c_uint32 * MAX_BINS), ("bin_window_lanes", C.c_uint32 * MAX_BINS),
                ("ms_total", C.c_float), ("ms_flops", C.c_float), ("ms_symbolic", C.c_float),
                ("ms_scan", C.c_float), ("ms_numeric", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
                ("n_launches", C.c_uint32), ("n_recorded", C.c_uint32),
                ("launches", Launch * MAX_LAUNCHES)]


class StreamStats(C.Structure):
    _fields_ = [("panels", C.c_uint64), ("products", C.c_uint64), ("nnz_c", C.c_uint64), ("max_panel_products", C.c_uint64),
                ("ms_total", C.c_float)]


PANEL_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int64),
                         C.POINTER(C.c_int32), C.POINTER(C.c_double))

# every symbol include/spada_b200.h declares: name -> (restype, argtypes)
_vp, _vpp = C.c_void_p, C.POINTER(C.c_void_p)
_u64p = C.POINTER(C.c_uint64)
SYMBOLS = {
    "spada_b200_abi_version": (C.c_int, []),
    "spada_b200_last_error": (C.c_char_p, []),
    "spada_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "spada_b200_host_alloc": (C.c_int, [_vpp, C.c_size_t]),
    "spada_b200_host_free": (C.c_int, [_vp]),
    "spada_b200_create": (C.c_int, [C.POINTER(Opts), _vpp]),
    "spada_b200_destroy": (None, [_vp]),
    "spada_b200_set_stream": (C.c_int, [_vp, _vp]),
    "spada_b200_synchronize": (C.c_int, [_vp]),
    "spada_b200_trim": (C.c_int, [_vp]),
    "spada_b200_upload": (C.c_int, [_vp, C.POINTER(CsrView), _vpp]),
    "spada_b200_upload32": (C.c_int, [_vp, C.POINTER(CsrView32), _vpp]),
    "spada_b200_csr_wrap_device": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vpp]),
    "spada_b200_csr_prepare": (C.c_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "spada_b200_csr_set_one_shot": (C.c_int, [_vp]),
    "spada_b200_transpose": (C.c_int, [_vp, _vp, _vpp]),
    "spada_b200_csr_shape": (C.c_int, [_vp, _u64p, _u64p, _u64p]),
    "spada_b200_csr_device_ptrs": (C.c_int, [_vp, _vpp, _vpp, _vpp]),
    "spada_b200_csr_download32": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "spada_b200_csr_free": (None, [_vp]),
    "spada_b200_spgemm_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vpp]),
    "spada_b200_spgemm": (C.c_int, [_vp, C.POINTER(CsrView), C.POINTER(CsrView), _vpp]),
    "spada_b200_spgemm32": (C.c_int, [_vp, C.POINTER(CsrView32), C.POINTER(CsrView32), _vpp]),
    "spada_b200_flops": (C.c_int, [_vp, _vp, _vp, _u64p, _u64p]),
    "spada_b200_plan_shards": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _u64p]),
    "spada_b200_cbuf_create": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vpp]),
    "spada_b200_cbuf_export": (C.c_int, [_vp, _vp]),
    "spada_b200_cbuf_import": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, _vpp]),
    "spada_b200_cbuf_free": (None, [_vp]),
    "spada_b200_cbuf_device_ptrs": (C.c_int, [_vp, _vpp, _vpp, _vpp]),
    "spada_b200_cbuf_nnz": (C.c_int, [_vp, _u64p]),
    "spada_b200_cbuf_copy32": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "spada_b200_shard_begin": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _u64p, _vpp]),
    "spada_b200_shard_finish": (C.c_int, [_vp, _vpp, C.c_uint32, C.c_uint64, _vp, C.c_uint32, C.POINTER(Stats)]),
    "spada_b200_shard_abort": (None, [_vp]),
    "spada_b200_group_create": (C.c_int, [C.POINTER(Opts), C.c_uint32, _vpp]),
    "spada_b200_group_spgemm": (C.c_int, [_vp, C.POINTER(CsrView), C.POINTER(CsrView), _vpp]),
    "spada_b200_group_spgemm32": (C.c_int, [_vp, C.POINTER(CsrView32), C.POINTER(CsrView32), _vpp]),
    "spada_b200_group_destroy": (None, [_vp]),
    "spada_b200_spgemm_stream": (C.c_int, [_vp, _vp, _vp, C.c_uint64, PANEL_SINK, _vp, C.POINTER(StreamStats)]),
    "spada_b200_spgemm_to_host": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_double), C.c_uint64, C.POINTER(StreamStats)]),
    "spada_b200_spgemm32_host_to_host": (C.c_int, [_vp, C.POINTER(CsrView32), C.POINTER(CsrView32), C.POINTER(C.c_int64),
                                                   C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_uint64,
                                                   C.POINTER(StreamStats)]),
    "spada_b200_result_shape": (C.c_int, [_vp, _u64p, _u64p, _u64p]),
    "spada_b200_result_copy": (C.c_int, [_vp, _u64p, _u64p, C.POINTER(C.c_double)]),
    "spada_b200_result_copy32": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "spada_b200_result_device_ptrs": (C.c_int, [_vp, _vpp, _vpp, _vpp]),
    "spada_b200_result_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "spada_b200_result_free": (None, [_vp]),
}

_lib = None


class SpadaB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS.get(code, code)}: {message}")
        self.code = code
        self.status = STATUS.get(code, str(code))


def lib():
    """Load libspada_b200.so; raise if it has not been built (never falls back to CPU code)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C spada-sim_b200/csrc).  This package has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if l.spada_b200_abi_version() != ABI_VERSION:
            raise ImportError(f"{LIB_PATH} has ABI version {l.spada_b200_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise SpadaB200Error(rc, lib().spada_b200_last_error().decode("utf-8", "replace"))
