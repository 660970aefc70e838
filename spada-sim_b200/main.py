"""``spada-sim`` CLI drop-in -- mirror of the reference's src/main.rs:30-121.

    spada-sim <simulator> <accelerator> <category> <workload> <configuration> [-p]

Same argv grammar, same loaders, same stdout sections; the simulated-hardware counters are
replaced by analytic element counts (see simulator.py).  Only ``AccurateSimu`` is implemented,
like upstream (main.rs:119 panics for the other modes).
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .frontend import parse_args, parse_config
from .gemm import GEMM
from .py2rust import load_mm_mat, load_pickled_gemms
from .simulator import Simulator
from .storage import CsrMatStorage, sort_by_length


def fnv1a64(data: bytes, h: int = 14695981039346656037) -> int:
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def dump_result(path: str, indptr, indices, data, n_cols: int) -> str:
    """Result sink (the reference prints ten rows of C and drops the rest, main.rs:113-116):
    all of C as a Matrix Market coordinate file (1-based, %.17g) + the digest line the compiled host prints."""
    indptr = np.asarray(indptr, dtype="<u8"); indices = np.asarray(indices, dtype="<u8"); data = np.asarray(data, dtype="<f8")
    rows = np.repeat(np.arange(len(indptr) - 1, dtype=np.int64), np.diff(indptr.astype(np.int64)))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (len(indptr) - 1, n_cols, len(data)))
        for r, c, v in zip(rows, indices, data):
            f.write("%d %d %.17g\n" % (r + 1, int(c) + 1, v))
    h = fnv1a64(data.tobytes(), fnv1a64(indices.tobytes(), fnv1a64(indptr.tobytes())))
    total = 0.0
    for v in data:            # same left-to-right sum as the compiled host
        total += float(v)
    return "C dumped: nnz %d sum %.17g fnv1a64 %016x" % (len(data), total, h)


def main(argv=None) -> int:
    cli = parse_args(argv)
    cfg = parse_config(cli.configuration)
    if cli.category == "NN":
        gemm = load_pickled_gemms(cfg.nn_filepath, cli.workload)
    else:
        gemm = GEMM.from_mat(cli.workload, load_mm_mat(cfg.ss_filepath, cli.workload))

    a_avg = gemm.a.nnz // gemm.a.shape[0]   # m == 0 divides by zero upstream too (main.rs:44)
    b_avg = gemm.b.nnz // gemm.b.shape[0]
    print(f"Get GEMM {gemm.name}")
    print(f"{gemm}")
    print(f"Avg row len of A: {a_avg}, Avg row len of B: {b_avg}")

    if cli.simulator != "AccurateSimu":
        raise SystemExit(f"Unimplemented simulator {cli.simulator}")

    dram_a, dram_b = CsrMatStorage.init_with_gemm(gemm)
    if cli.preprocess:
        dram_a.reorder_row(sort_by_length(dram_a))
    output_base_addr = len(dram_b.indptr)
    if cli.accelerator == "Op":
        block_shape = [cfg.lane_num, 1]
    else:
        block_shape = list(cfg.block_shape)

    sim = Simulator(cfg.pe_num, cfg.at_num, cfg.lane_num, cfg.cache_size, cfg.word_byte, output_base_addr,
                    block_shape, dram_a, dram_b, None, cli.accelerator, cfg.mem_latency, cfg.cache_latency,
                    cfg.freq, cfg.channel, cfg.bandwidth_per_channel)
    sim.execute()
    result = sim.get_exec_result()
    if os.environ.get("SPADA_B200_DUMP_C"):     # result sink: all of C + digest (stderr keeps stdout = the reference's)
        ip, ix, dx = sim.get_exec_csr()
        print(dump_result(os.environ["SPADA_B200_DUMP_C"], ip, ix, dx, gemm.b.shape[1]), file=sys.stderr)
    a_count, b_count, c_count = sim.get_a_mat_stat(), sim.get_b_mat_stat(), sim.get_c_mat_stat()
    cache_count = sim.get_cache_stat()

    print("-----Result-----")
    print("-----Access count")
    print(f"Execution count: {sim.get_exec_cycle()}")
    print(f"A matrix count: read {a_count[0]} write {a_count[1]}")
    print(f"B matrix count: read {b_count[0]} write {b_count[1]}")
    print(f"C matrix count: read {c_count[0]} write {c_count[1]}")
    print(f"Cache count: read {cache_count[0]} write {cache_count[1]}")
    print("-----Output product matrix")
    for row in result[:10]:
        print(row)
    return 0


if __name__ == "__main__":
    sys.exit(main())
