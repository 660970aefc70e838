#!/usr/bin/env python
"""bench.py -- SpGEMM GFLOP/s (2 x intermediate products / time) and % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rect|poisson|er|rmat|cari]
    python bench.py --impl reference ...     # the CPU restatement timed on the host cores

One "step" = one pass of the hot path (flop count + binning, sort passes, long rows, scan, placement) over one
synthetic operand pair that is already resident in HBM.  Default workload: BASELINE.json
configs[4], the rectangular power-law 1M x 4M matrix, A x A^T -- the configuration the metric's
"1/2/4/8 B200" clause is quoted on; it fits one GPU, so the same
workload is used at every N (strong scaling: A row-sharded by equal product count, B replicated,
C shards all-gathered).  Operands (0.8 GB) and C (3 GB) are far larger than the 126 MB L2, so
no explicit flush is needed between steps.

Prints ONE JSON line (rank 0).  `value` = 2*products*K / max-over-ranks device time;
`e2e` = same metric through the host-level C-ABI call (pinned host operands in, C copied back
to pinned host memory, both copies inside the timed region); `roofline` is for the dominant
kernel launch, timed with CUDA events on the stream the engine launches on; `cpu_baseline` is
the oracle (a port -- the Rust reference cannot be built here) on a bounded sample of rows.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_DESC = {
    "rect": "rect power-law 1M x 4M (avg 32 nnz/row, seed 2024), A x A^T  [BASELINE configs[4]]",
    "poisson": "2D Poisson 5-point 2048x2048 grid, A x A  [BASELINE configs[1]]",
    "er": "Erdos-Renyi 2M x 2M, 16 nnz/row (seed 1234), A x A  [BASELINE configs[2]]",
    "rmat": "R-MAT scale 21, edge factor 16, (0.45,0.22,0.22,0.11) (seed 42), A x A  [BASELINE configs[3]]",
    "cari": "matrices/cari.mtx (400 x 1200), A x A^T  [BASELINE configs[0]]",
}


def load_workload(pkg, name, scale):
    if name == "cari":
        import scipy.sparse as sp
        z = np.load(os.path.join(ROOT, "tests", "golden", "cari_csr.npz"))
        a = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        g = pkg.GEMM.from_mat("cari", a)
        return g.a, g.b
    a, b = pkg.workloads.build(name, scale)
    if scale >= 1.0 and name in pkg.workloads.KNOWN:
        assert a.nnz == pkg.workloads.KNOWN[name][2], "generator drifted from SURVEY.md 8d known answers"
    return a, b


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: an in-process NVML thread
    (pynvml, 10 ms period -- a 0.3 s timed region still gets ~30 samples); `nvidia-smi -lms` as the
    fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        import threading
        self.samples = []          # (wall time, sm MHz, reasons bitmask)
        self.smmax = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.p = None
        self.f = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self._start_smi(index)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.time(), mhz, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def _start_smi(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu=timestamp,{self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(1.0)   # nvidia-smi needs a moment before its first sample
        except OSError:
            self.p = None

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nv
            names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            win = [x for x in self.samples if t0 <= x[0] <= t1] or [x for x in self.samples if t0 - 0.05 <= x[0] <= t1 + 0.05]
            reasons = set()
            for _, _, rs in win:
                for n_, bit in names.items():
                    if rs & bit:
                        reasons.add(n_)
            if win:
                out.update(sm_mhz=statistics.median(x[1] for x in win), sm_max_mhz=self.smmax, reasons=sorted(reasons),
                           samples=len(win), source="nvml")
            try:
                nv.nvmlShutdown()
            except Exception:
                pass
            return out
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        import datetime
        sm, reasons, smmax = [], set(), None
        for r in rows:
            if len(r) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if not ((t0 - 0.05) <= ts <= (t1 + 0.05)):
                    continue
                sm.append(float(r[1])); smmax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
            except ValueError:
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=smmax, reasons=sorted(reasons), samples=len(sm),
                       source="nvidia-smi")
        return out


def cpu_oracle_rate(a, b, target_seconds, threads=None):
    """Oracle (CPU port) on a bounded sample: the first rows of A whose products sum to what the host
    can do in ~target_seconds.  Returns (gflops, cores, sample description, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    threads = threads or oracle.max_threads()
    lens_b = np.diff(b.indptr).astype(np.int64)
    per_row = np.add.reduceat(np.append(lens_b[a.indices], 0), a.indptr[:-1].astype(np.int64))
    per_row[np.diff(a.indptr) == 0] = 0
    csum = np.cumsum(per_row)
    total = int(csum[-1])
    # calibration on ~2M products
    r_cal = max(1, int(np.searchsorted(csum, min(total, 2_000_000), side="left")) + 1)
    t = time.perf_counter(); oracle.spgemm(a[:r_cal], b, threads=threads); dt = time.perf_counter() - t
    rate = max(csum[r_cal - 1], 1) / max(dt, 1e-6)
    want = int(min(total, rate * target_seconds))
    r = max(r_cal, int(np.searchsorted(csum, want, side="left")) + 1)
    r = min(r, a.shape[0])
    sub = a[:r]
    t = time.perf_counter(); oracle.spgemm(sub, b, threads=threads); dt = time.perf_counter() - t
    prods = int(csum[r - 1])
    return 2.0 * prods / dt / 1e9, threads, f"rows [0,{r}) of A = {prods} of {total} products", dt


def workload_config(pkg, args, a, b, products=None, nnz_c=None):
    """The `config` object both arms print: same keys and values for the same workload."""
    m, k, n = a.shape[0], a.shape[1], b.shape[1]
    known = pkg.workloads.KNOWN.get(args.workload) if args.scale >= 1.0 else None
    if known:
        products = products if products is not None else known[3]
        nnz_c = nnz_c if nnz_c is not None else known[4]
    alg = pkg.workloads.algorithmic_bytes(a.nnz, m, b.nnz, k, nnz_c) if nnz_c else None
    return {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale, "m": m, "k": k, "n": n, "nnz_a": int(a.nnz),
            "nnz_b": int(b.nnz), "products": products, "nnz_c": nnz_c,
            "l2": ("inputs+output larger than L2 (no flush)" if alg and alg > 4 * 126e6 else "working set near L2 size")}


def run_reference(args, pkg):
    """--impl reference: the reference's CPU implementation of the path.  The Rust simulator cannot
    be built in this image (no cargo, un-vendored crates), so this times the oracle port with every
    host core (set explicitly: torchrun exports OMP_NUM_THREADS=1), one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    a, b = load_workload(pkg, args.workload, args.scale)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    lens_b = np.diff(b.indptr).astype(np.int64)
    per_row = np.add.reduceat(np.append(lens_b[a.indices], 0), a.indptr[:-1].astype(np.int64))
    per_row[np.diff(a.indptr) == 0] = 0
    csum = np.cumsum(per_row)
    total = int(csum[-1])
    r_cal = max(1, int(np.searchsorted(csum, min(total, 2_000_000), side="left")) + 1)
    t = time.perf_counter(); oracle.spgemm(a[:r_cal], b, threads=threads); dt = time.perf_counter() - t
    rate = max(csum[r_cal - 1], 1) / max(dt, 1e-6)
    budget = 120.0 / max(1, args.steps + args.warmup)       # whole run within a couple of minutes
    want = int(min(total, rate * min(budget, 10.0)))
    r = min(a.shape[0], max(r_cal, int(np.searchsorted(csum, want, side="left")) + 1))
    sub = a[:r]
    prods = int(csum[r - 1])
    for _ in range(args.warmup):
        oracle.spgemm(sub, b, threads=threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        oracle.spgemm(sub, b, threads=threads)
    dt = time.perf_counter() - t
    v = 2.0 * prods * args.steps / dt / 1e9
    sample = f"rows [0,{r}) of A = {prods} of {total} products per step"
    cfg = workload_config(pkg, args, a, b, products=total if args.scale < 1.0 else None)
    print(json.dumps({
        "impl": "reference", "metric": "SpGEMM GFLOP/s (2 x intermediate products / time)", "value": v,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def pinned_copy(pkg, arr):
    """Copy a numpy array into cudaMallocHost memory; returns (numpy view, raw pointer)."""
    lib = pkg._abi.lib()
    p = C.c_void_p()
    pkg._abi.check(lib.spada_b200_host_alloc(C.byref(p), max(arr.nbytes, 1)))
    buf = (C.c_char * max(arr.nbytes, 1)).from_address(p.value)
    v = np.frombuffer(buf, dtype=arr.dtype, count=arr.size)
    v[...] = arr.ravel()
    return v, p


def add_launches(per_launch, stats):
    for L in stats["launches"]:
        d = per_launch.setdefault(L["name"], {"ms": 0.0, "n": 0, "products": 0, "rows": L["rows"], "grid": L["grid"]})
        d["ms"] += L["ms"]; d["n"] += 1; d["products"] += L["products"]


NVLINK_PEAK_GBS = 770.0   # measured peer-copy bandwidth per direction, /opt/skills/guides/B200_PROFILING.md
PLACEMENT_KERNELS = ("copy_rows", "copy_gather", "row_ptr_gather")   # move finished rows: no algorithmic work


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rect", choices=list(WORKLOAD_DESC))
    ap.add_argument("--scale", type=float, default=1.0, help="<1 shrinks the workload (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--two-phase", action="store_true", help="force the scratch-row pass for every row (exact-size C)")
    ap.add_argument("--single-pass", action="store_true", help="force the fused single pass for rows <= 512 products")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = the kernel that places C's rows stores them into every rank's buffers over NVLink "
                         "(the product); 'nccl' = separate NCCL all-gather after the compute (the baseline it replaces)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    pkg = importlib.import_module("spada-sim_b200")
    if args.impl == "reference":
        run_reference(args, pkg)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    D = importlib.import_module("spada-sim_b200.distributed")

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn):
        t0, t1 = ev(), ev()
        t0.record()
        out = fn()
        t1.record()
        torch.cuda.synchronize()
        return out, t0.elapsed_time(t1)

    stream = torch.cuda.current_stream()
    eng = pkg.Engine(device=local_rank, validate=True, stream=stream.cuda_stream, two_phase=args.two_phase,
                     single_pass=args.single_pass)

    # ---- operands resident in HBM (outside the timed steps; reported in `setup`) ------------------------------
    a = b = None
    if rank == 0:
        a, b = load_workload(pkg, args.workload, args.scale)
    setup = {}
    pg = None
    if world == 1:
        da = eng.upload(a)
        db = da if b is a else eng.upload(b)
        setup["prepare_ms"] = db.prepare()                   # fiber store of B (first use as B would build it anyway)
        lo, hi = 0, a.shape[0]
        dims = (a.shape[0], a.shape[1], b.shape[1], a.nnz, b.nnz)
        same_operand = b is a
    else:
        D.broadcast_csr(eng, a, device)                      # warm-up of the communicator
        (da, _ka), bc_a = timed(lambda: D.broadcast_csr(eng, a, device))   # NCCL broadcast over NVLink
        same = [b is a] if rank == 0 else [None]
        dist.broadcast_object_list(same, src=0)
        same_operand = same[0]
        bc_b = 0.0
        if same_operand:
            db, _kb = da, _ka
        else:
            (db, _kb), bc_b = timed(lambda: D.broadcast_csr(eng, b, device))
        setup["broadcast_ms"] = {"A": bc_a, "B": bc_b,
                                 "note": "host arrays of rank 0 -> its GPU -> ncclBroadcast to every rank (SURVEY 8e); once per operand"}
        setup["prepare_ms"] = db.prepare()
        bounds, setup["plan_ms"] = timed(lambda: D.plan_bounds(eng, da, db, world, device))
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        dims = (da.shape[0], da.shape[1], db.shape[1], da.nnz, db.nnz)
    m, k, n, nnz_a, nnz_b = dims

    # GEMM::from_mat's transpose (gemm.rs:41-53) on the device, timed once for the record (not part of a step:
    # the workload builder already holds B = A^T on the host, like the reference after its loader)
    if world == 1 and not same_operand and a.shape[0] != a.shape[1]:
        eng.transpose(da).free()                             # warm-up (pool blocks)
        dt, setup["device_transpose_ms"] = timed(lambda: eng.transpose(da))
        assert dt.shape == db.shape and dt.nnz == db.nnz
        dt.free()

    if world > 1:
        cap = torch.zeros(1, dtype=torch.int64, device=device)
        if rank == 0:
            cap[0] = eng.flops(da, db)               # nnz(C) <= intermediate products
        dist.broadcast(cap, src=0)
        if args.gather == "peer":
            pg = D.PeerGather(eng, m, n, int(cap[0]), device)
        else:
            nccl_out = (torch.empty(m + 1, dtype=torch.int64, device=device),
                        torch.empty(int(cap[0]), dtype=torch.int32, device=device),
                        torch.empty(int(cap[0]), dtype=torch.float64, device=device))

    gathered = None

    def step(compute_only=False):
        """One pass of the hot path; returns the engine stats of this rank's part."""
        nonlocal gathered
        if world == 1:
            return eng.spgemm_dev(da, db, lo, hi).stats()
        if pg is not None:
            return pg.step(da, db, lo, hi, compute_only=compute_only)
        res = eng.spgemm_dev(da, db, lo, hi)
        if not compute_only:
            lp, lc, lv = D.result_views(res, device)
            gathered = D.allgather_csr(lp, lc, lv, out=nccl_out)
            torch.cuda.current_stream().synchronize()   # the Result's pool blocks may be reused once it is freed
        return res.stats()

    for _ in range(args.warmup):
        st = step()
    if world > 1 and pg is not None:
        nnz_c = pg.own.nnz
        products = int(cap[0])
    elif world > 1:
        nnz_c = int(gathered[0][-1])
        products = int(cap[0])
    else:
        products, nnz_c = st["products"], st["nnz_c"]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    launches = 0
    t_wall0 = time.time()
    e0.record()
    engine_ms = 0.0
    for _ in range(args.steps):
        st_k = step()
        launches += st_k["n_launches"]
        engine_ms += float(st_k.get("ms_total", 0.0))
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms[0])
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_per_step = total_ms / args.steps
    gflops = 2.0 * products / (ms_per_step * 1e-3) / 1e9

    # ---- N > 1: the same steps without the exchange (compute only), NVLink bytes, and a check of the gathered C ----
    multi = None
    if world > 1:
        for _ in range(2):
            step(compute_only=True)
        torch.cuda.synchronize()
        dist.barrier()
        c0, c1 = ev(), ev()
        c0.record()
        for _ in range(5):
            step(compute_only=True)
        c1.record()
        torch.cuda.synchronize()
        cms = torch.tensor([c0.elapsed_time(c1) / 5], dtype=torch.float64, device=device)
        dist.all_reduce(cms, op=dist.ReduceOp.MAX)
        step()                                            # leave a complete C in the buffers
        torch.cuda.synchronize()
        dist.barrier()
        if pg is not None:
            g_ptr, g_col, g_val = pg.result()
            sent = torch.tensor([pg.nvlink_bytes() + 8 * (hi - lo + 1) * (world - 1)], dtype=torch.float64, device=device)
        else:
            g_ptr, g_col, g_val = gathered[0], gathered[1][:nnz_c], gathered[2][:nnz_c]
            sent = torch.tensor([0.0], dtype=torch.float64, device=device)
        # every rank compares its gathered C with the C it computes alone on one GPU (bit for bit, on the device)
        ref = eng.spgemm_dev(da, db)
        rp, rc, rv = D.result_views(ref, device)
        same_c = bool(torch.equal(rp, g_ptr[:m + 1]) and torch.equal(rc, g_col[:ref.nnz]) and
                      torch.equal(rv.view(torch.int64), g_val[:ref.nnz].view(torch.int64)))
        flag = torch.tensor([1 if same_c else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        mx = sent.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ref = rp = rc = rv = None
        multi = {"gather": args.gather, "compute_only_ms_per_step": float(cms[0]),
                 "gathered_c_identical_to_one_gpu_on_every_rank": bool(int(flag[0])),
                 "nvlink_bytes_sent_per_rank_per_step_max": float(mx[0]),
                 "nvlink_floor_ms": float(mx[0]) / (NVLINK_PEAK_GBS * 1e9) * 1e3,
                 "note": "value includes the exchange of C; compute_only = same steps storing only into the rank's own buffers"}

    # ---- N > 1 end to end: pinned host operands on rank 0 -> H2D -> NCCL broadcast -> shard plan -> sharded product with
    # the peer-store gather -> the whole C back in pinned host memory on rank 0; every step repeats all of it ---------
    e2e = None
    if world > 1 and pg is not None and args.e2e_steps > 0:
        pa = pb_ = None
        o_ptr = o_col = o_val = None
        if rank == 0:
            pa = D.PinnedCsr(a)
            pb_ = pa if same_operand else D.PinnedCsr(b)
            o_ptr = torch.empty(m + 1, dtype=torch.int64).pin_memory()
            o_col = torch.empty(nnz_c, dtype=torch.int32).pin_memory()
            o_val = torch.empty(nnz_c, dtype=torch.float64).pin_memory()

        def e2e_multi():
            xa, _k1 = D.broadcast_csr(eng, pa, device)
            xb, _k2 = (xa, _k1) if same_operand else D.broadcast_csr(eng, pb_, device)
            bnd = D.plan_bounds(eng, xa, xb, world, device)
            pg.step(xa, xb, int(bnd[rank]), int(bnd[rank + 1]))
            if rank == 0:
                g_ptr, g_col, g_val = pg.result()
                o_ptr.copy_(g_ptr, non_blocking=True)
                o_col.copy_(g_col[:nnz_c], non_blocking=True)
                o_val.copy_(g_val[:nnz_c], non_blocking=True)
            torch.cuda.synchronize()
            xa.free()
            if xb is not xa:
                xb.free()
        e2e_multi()
        dist.barrier()
        f0, f1 = ev(), ev()
        f0.record()
        for _ in range(args.e2e_steps):
            e2e_multi()
        f1.record()
        torch.cuda.synchronize()
        ems = torch.tensor([f0.elapsed_time(f1) / args.e2e_steps], dtype=torch.float64, device=device)
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert int(o_ptr[-1]) == nnz_c
            e2e = {"value": 2.0 * products / (float(ems[0]) * 1e-3) / 1e9, "unit": "GFLOP/s",
                   "h2d_bytes_per_step": (8 * (m + 1) + 12 * nnz_a) + (0 if same_operand else 8 * (k + 1) + 12 * nnz_b),
                   "d2h_bytes_per_step": 8 * (m + 1) + 12 * nnz_c, "ms_per_step": float(ems[0]), "steps": args.e2e_steps,
                   "note": "rank 0: pinned host CSR -> H2D -> NCCL broadcast of A and B -> shard plan -> sharded product with "
                           "peer-store gather -> whole C to rank 0's pinned host memory; max over ranks; one-shot operands "
                           "(no fiber store), like the N = 1 leg"}

    # ---- per-launch durations: the engine overlaps the long rows with the sort bins on a side stream, so the event
    # times of the timed region overlap too.  For the roofline every kernel is timed alone: a second handle with
    # SPADA_B200_FLAG_SERIAL over the same device arrays (this rank's rows), 3 warm-up + 5 recorded steps.
    if pg is not None:
        pg.close()
        pg = None
    if world > 1:
        nccl_out = gathered = g_ptr = g_col = g_val = None
    eng.trim()          # hand the first handle's cached blocks back: the second handle needs the same footprint
    ser = pkg.Engine(device=local_rank, validate=False, stream=stream.cuda_stream, two_phase=args.two_phase or world > 1,
                     single_pass=args.single_pass and world == 1, serial=True)
    sa = ser.wrap_device(da.shape[0], da.shape[1], da.nnz, *da.device_ptrs(), keepalive=da)
    sb = sa if db is da else ser.wrap_device(db.shape[0], db.shape[1], db.nnz, *db.device_ptrs(), keepalive=db)
    sb.prepare()
    per_launch = {}
    local_products = 0
    for it in range(8):
        r_ = ser.spgemm_dev(sa, sb, lo, hi)
        if it >= 3:
            s_ = r_.stats()
            local_products = s_["products"]
            add_launches(per_launch, s_)
        r_ = None
    sa.free()
    if sb is not sa:
        sb.free()
    ser.close()
    timing_note = ("launch times from 5 serialised steps after the timed region (SPADA_B200_FLAG_SERIAL, rank 0's rows): the "
                   "timed steps run the long rows on a side stream beside the sort bins")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant COMPUTE kernel (placement kernels do no algorithmic work) ---------------------
    peak, peak_src = peaks()
    alg_bytes = pkg.workloads.algorithmic_bytes(nnz_a, m, nnz_b, k, nnz_c)
    compute = {k_: v for k_, v in per_launch.items() if k_ not in PLACEMENT_KERNELS and v["products"] > 0}
    dom_name, dom = max((compute or per_launch).items(), key=lambda kv: kv[1]["ms"])
    dom_ms = dom["ms"] / dom["n"]
    # algorithmic bytes of one launch = whole-path compulsory bytes x the launch's share of ALL products of the workload
    share = (dom["products"] / dom["n"] / products) if products else 1.0
    dom_bytes = alg_bytes * share
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload, {}).get(dom_name.split("#")[0])   # "#k": wave k of the long rows
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": dom_name, "kernel_ms": dom_ms, "algorithmic_bytes": dom_bytes,
                "peak_source": peak_src, "timing": timing_note,
                "whole_path": {"algorithmic_bytes": alg_bytes, "ms": ms_per_step,
                               "achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                "launch_ms": {k_: v["ms"] / v["n"] for k_, v in per_launch.items()}}

    # ---- e2e: host-level C-ABI call with pinned host operands, C copied back (N = 1) --------------------------
    if world == 1 and args.e2e_steps > 0:
        e2e = e2e_single(pkg, eng, a, b, m, k, nnz_a, nnz_b, nnz_c, products, args.e2e_steps, torch)

    # ---- CPU baseline: the oracle port on a bounded sample of the same workload -------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, secs = cpu_oracle_rate(a, b, target_seconds=12.0)
        cpu = {"value": v, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample, "seconds": secs}

    cfg = workload_config(pkg, args, a, b, products=products, nnz_c=nnz_c)
    cfg["parallelism"] = (f"rows of A sharded over {world} GPU(s) by equal product count; B replicated (NCCL broadcast); "
                          + ("C gathered by peer stores from the placement kernel" if world > 1 and args.gather == "peer"
                             else "C all-gathered with NCCL" if world > 1 else "single GPU"))
    line = {
        "metric": "SpGEMM GFLOP/s (2 x intermediate products / time)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "setup": setup, "multi_gpu": multi, "bins": st["bins"],
        # cross-check of the timed region: every product ends with a sync of the engine's stream, so the events around
        # the K steps bracket all of their kernels; this is the engine's own CUDA-event span (first kernel of a product to
        # its last, on the stream the kernels run on), mean over this rank's timed steps
        "engine_ms_per_step": engine_ms / max(args.steps, 1),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def e2e_single(pkg, eng, a, b, m, k, nnz_a, nnz_b, nnz_c, products, steps, torch):
    """The same metric through the host-level C-ABI call: pinned host operands in, the whole C back in pinned host
    memory, both copies inside the timed region."""
    eng.synchronize()
    abi = pkg._abi
    lib = abi.lib()
    keep = []

    def view32(mat):
        ip, p1 = pinned_copy(pkg, np.ascontiguousarray(mat.indptr, dtype=np.int32))
        ix, p2 = pinned_copy(pkg, np.ascontiguousarray(mat.indices, dtype=np.int32))
        dx, p3 = pinned_copy(pkg, np.ascontiguousarray(mat.data, dtype=np.float64))
        keep.extend([p1, p2, p3])
        return abi.CsrView32(mat.shape[0], mat.shape[1], mat.nnz, ip.ctypes.data_as(C.POINTER(C.c_int32)),
                             ix.ctypes.data_as(C.POINTER(C.c_int32)), dx.ctypes.data_as(C.POINTER(C.c_double)))
    va = view32(a)
    vb = va if b is a else view32(b)
    o_ptr, q1 = pinned_copy(pkg, np.zeros(m + 1, dtype=np.int64))
    o_col, q2 = pinned_copy(pkg, np.zeros(nnz_c, dtype=np.int32))
    o_val, q3 = pinned_copy(pkg, np.zeros(nnz_c, dtype=np.float64))
    keep.extend([q1, q2, q3])
    h2d = (4 * (m + 1) + 12 * nnz_a) + (0 if b is a else 4 * (k + 1) + 12 * nnz_b)
    d2h = 8 * (m + 1) + 12 * nnz_c

    def e2e_step():
        # ONE ABI call: B up, then A up in row panels while the panels already there are computed and their results
        # come down (validation included)
        abi.check(lib.spada_b200_spgemm32_host_to_host(eng._h, C.byref(va), C.byref(vb),
                                                       o_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                       o_col.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       o_val.ctypes.data_as(C.POINTER(C.c_double)), len(o_col), None))
    e2e_step()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        e2e_step()
    f1.record()
    torch.cuda.synchronize()
    e2e_ms = f0.elapsed_time(f1) / steps
    assert int(o_ptr[-1]) == nnz_c
    out = {"value": 2.0 * products / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": steps,
           "note": "spada_b200_spgemm32_host_to_host: pinned host CSR in (B first, A in row panels on an upload stream), "
                   "validation, C computed panel by panel while the next panel's entries go up and the previous panel's "
                   "result comes down, whole C in pinned host arrays.  Operands of one product get no fiber store: B is "
                   "gathered through row_ptr (the resident-operand steps above build it once, outside the step: "
                   "setup.prepare_ms)"}
    for p in keep:
        lib.spada_b200_host_free(p)
    return out


if __name__ == "__main__":
    main()
