"""N>1 host logic on CPU: world_size 2 over gloo.  Each rank computes its row shard with the
oracle (the checker -- no GPU here), then the product's allgather_csr assembles the whole C."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_csr


def _worker(rank, world, port, q, shape=(301, 200, 150)):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    D = importlib.import_module("spada-sim_b200.distributed")
    m_, k_, n_ = shape
    try:
        a = random_csr(m_, k_, density=0.03, seed=21) if rank == 0 else None
        b = random_csr(k_, n_, density=0.05, seed=22) if rank == 0 else None
        _, (ap, aj, ax) = D.broadcast_csr(None, a, torch.device("cpu"))
        _, (bp, bj, bx) = D.broadcast_csr(None, b, torch.device("cpu"))
        import scipy.sparse as sp
        A = sp.csr_matrix((ax.numpy(), aj.numpy(), ap.numpy()), shape=(m_, k_))
        B = sp.csr_matrix((bx.numpy(), bj.numpy(), bp.numpy()), shape=(k_, n_))
        bounds = D.balanced_bounds(oracle.flops(A, B) + 1, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        cp, cj, cx = oracle.spgemm(A[lo:hi], B)
        gp, gj, gx = D.allgather_csr(torch.from_numpy(cp), torch.from_numpy(cj), torch.from_numpy(cx))
        fp, fj, fx = oracle.spgemm(A, B)
        ok = (np.array_equal(gp.numpy(), fp) and np.array_equal(gj.numpy(), fj) and np.array_equal(gx.numpy(), fx))
        q.put((rank, ok, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_allgather_csr_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    spans = sorted(s for _, _, s in res)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 301


def test_allgather_csr_world3_with_empty_shards():
    # three ranks, fewer rows than shards x waves: some shards (and whole waves of a rank) are empty
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 3, port, q, (7, 40, 30))) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    spans = sorted(s for _, _, s in res)
    assert spans[0][0] == 0 and spans[-1][1] == 7 and all(spans[i][1] == spans[i + 1][0] for i in range(2))
