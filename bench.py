#!/usr/bin/env python
"""bench.py -- SpGEMM GFLOP/s (2 x intermediate products / time) and % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rect|poisson|er|rmat|cari]
    python bench.py --impl reference ...     # the CPU restatement timed on the host cores

One "step" = one pass of the hot path (flop count + binning, symbolic, scan, numeric) over one
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


def run_reference(args, pkg):
    """--impl reference: the reference's CPU implementation of the path.  The Rust simulator cannot
    be built in this image (no cargo, un-vendored crates), so this times the oracle port with every
    host thread, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    a, b = load_workload(pkg, args.workload, args.scale)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    threads = oracle.max_threads()
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
    print(json.dumps({
        "impl": "reference", "metric": "SpGEMM GFLOP/s (2 x intermediate products / time)", "value": v,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale},
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
    ap.add_argument("--two-phase", action="store_true", help="force separate symbolic/numeric passes (exact-size C)")
    ap.add_argument("--single-pass", action="store_true", help="force the fused single pass for rows <= 512 products")
    ap.add_argument("--gather-waves", type=int, default=1,
                    help="N > 1: row waves per rank; with more than one, the all-gather of wave j overlaps the computation "
                         "of wave j+1 (measured at N=2 on rect: 6.9 / 8.7 / 9.3 ms per step for 1 / 2 / 3 waves -- the NCCL "
                         "copy kernels queue behind the SpGEMM kernels and every wave adds host round trips, so 1 is the default)")
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

    stream = torch.cuda.current_stream()
    eng = pkg.Engine(device=local_rank, validate=True, stream=stream.cuda_stream, two_phase=args.two_phase,
                     single_pass=args.single_pass)

    # ---- operands resident in HBM (not timed) ---------------------------------------------------
    a = b = None
    if rank == 0:
        a, b = load_workload(pkg, args.workload, args.scale)
    if world == 1:
        da = eng.upload(a)
        db = da if b is a else eng.upload(b)
        lo, hi = 0, a.shape[0]
        dims = (a.shape[0], a.shape[1], b.shape[1], a.nnz, b.nnz)
    else:
        da, _ka = D.broadcast_csr(eng, a, device)            # NCCL broadcast over NVLink
        same = [b is a] if rank == 0 else [None]
        dist.broadcast_object_list(same, src=0)
        db, _kb = (da, _ka) if same[0] else D.broadcast_csr(eng, b, device)
        db.prepare()                                         # fiber store of the replicated B (not timed, like the broadcast)
        waves = max(1, args.gather_waves)
        bounds = D.plan_bounds(eng, da, db, world * waves, device)   # shard j * world + r = wave j of rank r
        lo, hi = 0, 0
        dims = (da.shape[0], da.shape[1], db.shape[1], da.nnz, db.nnz)
    m, k, n, nnz_a, nnz_b = dims

    # GEMM::from_mat's transpose (gemm.rs:41-53) on the device, timed once for the record (not part of a step:
    # the workload builder already holds B = A^T on the host, like the reference after its loader)
    transpose_ms = None
    if world == 1 and b is not a and a.shape[0] != a.shape[1]:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.transpose(da).free()                             # warm-up (pool blocks)
        t0.record()
        dt = eng.transpose(da)
        t1.record()
        torch.cuda.synchronize()
        assert dt.shape == db.shape and dt.nnz == db.nnz
        transpose_ms = t0.elapsed_time(t1)
        dt.free()

    gathered = None
    wave_gather = None
    if world > 1:
        cap = torch.zeros(1, dtype=torch.int64, device=device)
        if rank == 0:
            cap[0] = eng.flops(da, db)               # nnz(C) <= intermediate products
        dist.broadcast(cap, src=0)
        wave_gather = D.WaveGather(m, int(cap[0]), device)

    def step():
        """One pass of the hot path; returns the stats of the engine calls it made."""
        nonlocal gathered
        if world == 1:
            return [eng.spgemm_dev(da, db, lo, hi).stats()]
        wave_gather.reset()
        stats = []
        for j in range(waves):
            s_ = j * world + rank
            res = eng.spgemm_dev(da, db, int(bounds[s_]), int(bounds[s_ + 1]))
            stats.append(res.stats())
            lp, lc, lv = D.result_views(res, device)
            wave_gather.add(lp, lc, lv, keepalive=res)       # asynchronous: wave j travels while wave j+1 is computed
        gathered = wave_gather.finish()
        return stats

    for _ in range(args.warmup):
        sts = step()
    tot = torch.tensor([sum(x["products"] for x in sts), sum(x["nnz_c"] for x in sts)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(tot)
    products, nnz_c = int(tot[0]), int(tot[1])
    st0 = {"products": sum(x["products"] for x in sts), "bins": sts[0]["bins"]}

    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    per_launch = {}
    compute_ms = 0.0
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        for st in step():                     # host-side copies of event timings already taken by the engine
            launches += st["n_launches"]
            compute_ms += st["ms_total"]
            for L in st["launches"]:
                d = per_launch.setdefault(L["name"], {"ms": 0.0, "n": 0, "products": 0, "rows": L["rows"], "grid": L["grid"]})
                d["ms"] += L["ms"]; d["n"] += 1; d["products"] += L["products"]
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1), compute_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms, compute_ms = float(ms[0]), float(ms[1])
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_per_step = total_ms / args.steps
    gflops = 2.0 * products / (ms_per_step * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-launch durations: the engine overlaps the heavy / huge bins with the sort bins on a side stream, so the
    # event times of the timed region overlap too.  For the roofline every kernel is timed alone: a second handle with
    # SPADA_B200_FLAG_SERIAL over the same device arrays, 3 warm-up + 5 recorded steps, CUDA events on its stream.
    timing_note = "launch times from the timed region"
    if world == 1:
        eng.trim()          # hand the first handle's cached blocks back: the second handle needs the same footprint
        ser = pkg.Engine(device=local_rank, validate=False, stream=stream.cuda_stream, two_phase=args.two_phase,
                         single_pass=args.single_pass, serial=True)
        sa = ser.wrap_device(da.shape[0], da.shape[1], da.nnz, *da.device_ptrs(), keepalive=da)
        sb = sa if db is da else ser.wrap_device(db.shape[0], db.shape[1], db.nnz, *db.device_ptrs(), keepalive=db)
        sb.prepare()
        per_launch = {}
        for it in range(8):
            r_ = ser.spgemm_dev(sa, sb, lo, hi)
            if it >= 3:
                for L in r_.stats()["launches"]:
                    d = per_launch.setdefault(L["name"], {"ms": 0.0, "n": 0, "products": 0, "rows": L["rows"],
                                                          "grid": L["grid"]})
                    d["ms"] += L["ms"]; d["n"] += 1; d["products"] += L["products"]
            r_ = None
        sa.free()
        if sb is not sa:
            sb.free()
        ser.close()
        timing_note = ("launch times from 5 serialised steps after the timed region (SPADA_B200_FLAG_SERIAL): the timed "
                       "steps run the heavy/huge bins on a side stream beside the sort bins")

    # ---- roofline of the dominant kernel launch ----------------------------------------------------
    peak, peak_src = peaks()
    alg_bytes = pkg.workloads.algorithmic_bytes(nnz_a, m, nnz_b, k, nnz_c)
    dom_name, dom = max(per_launch.items(), key=lambda kv: kv[1]["ms"])
    dom_ms = dom["ms"] / dom["n"]
    # algorithmic bytes of one launch = whole-path compulsory bytes x the launch's share of products
    local_products = st0["products"]
    share = (dom["products"] / dom["n"] / local_products) if local_products else 1.0   # products of ONE launch
    if dom["products"] == 0:
        share = 1.0
    dom_bytes = alg_bytes / world * share
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload, {}).get(dom_name)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": dom_name, "kernel_ms": dom_ms, "algorithmic_bytes": dom_bytes,
                "peak_source": peak_src, "timing": timing_note,
                "whole_path": {"algorithmic_bytes": alg_bytes, "ms": ms_per_step,
                               "achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                "launch_ms": {k_: v["ms"] / v["n"] for k_, v in per_launch.items()}}

    # ---- e2e: host-level C-ABI call with pinned host operands, C copied back (rank 0, N = 1) ---------
    e2e = None
    if world == 1 and args.e2e_steps > 0:
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
            out = C.c_void_p()
            abi.check(lib.spada_b200_spgemm32(eng._h, C.byref(va), C.byref(vb), C.byref(out)))
            abi.check(lib.spada_b200_result_copy32(out, o_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   o_col.ctypes.data_as(C.POINTER(C.c_int32)),
                                                   o_val.ctypes.data_as(C.POINTER(C.c_double))))
            lib.spada_b200_result_free(out)
        e2e_step()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        f1.record()
        torch.cuda.synchronize()
        e2e_ms = f0.elapsed_time(f1) / args.e2e_steps
        assert int(o_ptr[-1]) == nnz_c
        e2e = {"value": 2.0 * products / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": args.e2e_steps,
               "note": "spada_b200_spgemm32 + result_copy32: pinned host CSR in, validation, compute, whole C to pinned host"}
        for p in keep:
            lib.spada_b200_host_free(p)

    # ---- CPU baseline: the oracle port on a bounded sample of the same workload -------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, secs = cpu_oracle_rate(a, b, target_seconds=12.0)
        cpu = {"value": v, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample, "seconds": secs}

    line = {
        "metric": "SpGEMM GFLOP/s (2 x intermediate products / time)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "scale": args.scale, "m": m, "k": k, "n": n,
                   "nnz_a": nnz_a, "nnz_b": nnz_b, "products": products, "nnz_c": nnz_c, "device_transpose_ms": transpose_ms,
                   "l2": "inputs+output larger than L2 (no flush)" if alg_bytes > 4 * 126e6 else "working set near L2 size",
                   "parallelism": f"rows of A sharded over {world} GPU(s) by equal product count; B replicated; C all-gathered"
                                  + (f" in {waves} waves overlapped with the computation" if world > 1 and waves > 1 else "")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "compute_only": {"ms_per_step": compute_ms / args.steps,
                         "value": 2.0 * products / (compute_ms / args.steps * 1e-3) / 1e9 if compute_ms else None,
                         "note": "engine device time per step (max over ranks), without the C all-gather"},
        "bins": st0["bins"],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
