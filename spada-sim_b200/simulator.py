"""``Simulator`` -- the reference's operator interface for the hot path, backed by the engine.

Mirrors the triple ``main`` calls (main.rs:74-100): ``Simulator::new`` (simulator.rs:431-448,
same positional arguments), ``execute`` (:509) and ``get_exec_result -> Vec<CsrRow>`` (:1034),
plus the stat getters (:1008-1032).  ``execute`` runs C = A x B on the B200 through the C ABI;
the cycle / traffic model behind the reference's counters is out of scope, so the getters
return analytic element counts with the reference's own accounting rules:
  A read  = 2 per stored nonzero fetched   (storage.rs:313-315)
  B read  = 2 per streamed B element = 2 x intermediate products
  C write = 2 per output nonzero + 1 per row (storage.rs:201-203)
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

import os

from . import _abi
from .engine import Engine, Group, Result
from .storage import CsrMatStorage, CsrRow

import ctypes as C


class Simulator:
    def __init__(self, pe_num: int, at_num: int, lane_num: int, cache_size: int, word_byte: int,
                 output_base_addr: int, default_block_shape, a_matrix: CsrMatStorage, b_matrix: CsrMatStorage,
                 psum_matrix, accelerator: str, mem_latency: int = 0, cache_latency: int = 0, freq: float = 1.0,
                 channel: int = 1, bandwidth_per_channel: float = 1.0, device: int = -1):
        self.a_matrix = a_matrix
        self.b_matrix = b_matrix
        self.accelerator = accelerator
        self.lane_num = lane_num
        self.default_block_shape = list(default_block_shape)
        # SPADA_B200_GPUS=N: every product sharded over N GPUs of this process (spada_b200_group_*)
        self.n_gpus = max(1, int(os.environ.get("SPADA_B200_GPUS", "1")))
        if self.n_gpus > 1:
            self.engine = Group(self.n_gpus, accelerator=accelerator, lane_num=lane_num, block_shape=default_block_shape)
        else:
            self.engine = Engine(device=device, accelerator=accelerator, lane_num=lane_num,
                                 block_shape=default_block_shape)
        self._result: Optional[Result] = None
        self._stats = None

    def _view(self, m: CsrMatStorage):
        cols, rows = m.mat_shape
        v = _abi.CsrView(rows, cols, len(m.data), m.indptr.ctypes.data_as(C.POINTER(C.c_uint64)),
                         m.indices.ctypes.data_as(C.POINTER(C.c_uint64)), m.data.ctypes.data_as(C.POINTER(C.c_double)))
        return v

    def execute(self) -> None:
        """simulator.rs:509-890 -- here: one spada_b200_spgemm call over the usize/f64 buffers."""
        out = C.c_void_p()
        va = self._view(self.a_matrix)
        same = (self.b_matrix.indptr is self.a_matrix.indptr and self.b_matrix.indices is self.a_matrix.indices
                and self.b_matrix.data is self.a_matrix.data)
        vb = va if same else self._view(self.b_matrix)
        call = _abi.lib().spada_b200_group_spgemm if self.n_gpus > 1 else _abi.lib().spada_b200_spgemm
        _abi.check(call(self.engine._h, C.byref(va), C.byref(vb), C.byref(out)))
        self._result = Result(self.engine, out)
        self._stats = self._result.stats()

    def get_exec_result(self) -> List[CsrRow]:
        """simulator.rs:1034-1062: one CsrRow per A row, raw row order, empty rows kept."""
        ip, ix, dx = self._result.to_host_usize()
        ip = ip.astype(np.int64)
        return [CsrRow.new_from_data(r, dx[ip[r]:ip[r + 1]], ix[ip[r]:ip[r + 1]]) for r in range(len(ip) - 1)]

    def get_exec_csr(self):
        """Whole C as (indptr u64, indices u64, data f64) without building per-row objects."""
        return self._result.to_host_usize()

    # stat getters (simulator.rs:1008-1032)
    def get_exec_cycle(self) -> int:
        return 0  # cycle model out of scope

    def get_a_mat_stat(self):
        return [2 * int(self._stats["nnz_a"]), 0]

    def get_b_mat_stat(self):
        return [2 * int(self._stats["products"]), 0]

    def get_c_mat_stat(self):
        return [0, 2 * int(self._stats["nnz_c"]) + int(self._stats["rows"])]

    def get_cache_stat(self):
        return [0, 0]

    def engine_stats(self) -> dict:
        return self._stats
