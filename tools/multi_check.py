#!/usr/bin/env python
"""Hardware check of the sharded path (run under torchrun, one rank per GPU, or with --group in one process).

Every rank computes the whole C by itself (one GPU, the path the parity tests pin against the oracle) and compares it
bit for bit with the C that the sharded product gathered into its buffers: row_ptr, col_idx and values.
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def operands(pkg, name, scale):
    from conftest import random_csr
    if name == "mixed":
        rng = np.random.default_rng(5)
        lens = rng.choice([0, 1, 3, 8, 20, 60, 150, 400, 900], size=3000, p=[.1, .2, .2, .2, .14, .1, .04, .01, .01])
        a = random_csr(3000, 2000, row_nnz=lens, seed=6, values="signed")
        b = random_csr(2000, 40000, row_nnz=rng.choice([0, 2, 9, 30, 200], size=2000), seed=7, values="signed")
        return a, b
    return pkg.workloads.build(name, scale)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mixed")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--group", type=int, default=0, help="single process: spada_b200_group over this many GPUs")
    args = ap.parse_args()
    pkg = importlib.import_module("spada-sim_b200")
    if args.group:
        a, b = operands(pkg, args.workload, args.scale)
        e = pkg.Engine(device=0)
        ref = e.spgemm(a, b).to_host()
        e.close()
        g = pkg.Group(args.group)
        for it in range(2):
            r = g.spgemm(a, b)
            got = r.to_host()
            ok = all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(got, ref))
            print(f"group of {args.group}: call {it}: nnz {r.nnz}, identical to one GPU: {ok}", flush=True)
            assert ok
            r.free()
        g.close()
        return
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D = importlib.import_module("spada-sim_b200.distributed")
    eng = pkg.Engine(device=local, stream=torch.cuda.current_stream().cuda_stream)
    a = b = None
    if rank == 0:
        a, b = operands(pkg, args.workload, args.scale)
    da, _ka = D.broadcast_csr(eng, a, dev)
    same = [b is a] if rank == 0 else [None]
    dist.broadcast_object_list(same, src=0)
    db, _kb = (da, _ka) if same[0] else D.broadcast_csr(eng, b, dev)
    db.prepare()
    bounds = D.plan_bounds(eng, da, db, world, dev)
    cap = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        cap[0] = eng.flops(da, db)
    dist.broadcast(cap, src=0)
    # a second product with the rows of A in reverse order: other shard bounds and other nnz per shard, so that a step
    # which picked up anything of the step before it (a stale nnz, a stale offset) cannot come out right
    a2 = a[::-1].tocsr() if rank == 0 else None
    da2, _ka2 = D.broadcast_csr(eng, a2, dev)
    bounds2 = D.plan_bounds(eng, da2, db, world, dev)
    pg = D.PeerGather(eng, da.shape[0], db.shape[1], int(cap[0]), dev)
    cases = [(da, bounds, eng.spgemm_dev(da, db).to_host()), (da2, bounds2, eng.spgemm_dev(da2, db).to_host())]
    ok = True
    for it in range(4):
        xa, bnd, ref = cases[it % 2]
        pg.step(xa, db, int(bnd[rank]), int(bnd[rank + 1]))
        torch.cuda.synchronize()
        got = pg.own.to_host()
        same_bits = all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(got, ref))
        ok = ok and same_bits
        print(f"rank {rank}/{world} step {it} ({'A' if it % 2 == 0 else 'A reversed'}): gathered nnz {len(got[1])}, "
              f"rows [{bnd[rank]}, {bnd[rank + 1]}), identical to one GPU: {same_bits}", flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    pg.close()
    dist.destroy_process_group()
    assert int(t[0]) == 1, "sharded C differs from the single-GPU C"


if __name__ == "__main__":
    main()
