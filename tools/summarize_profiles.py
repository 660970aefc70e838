#!/usr/bin/env python
"""Turns gpurun_out/ (scratch) into the tracked summaries under profiles/ for this round.

    python tools/summarize_profiles.py r02

Inputs (written by tools/gpu_final.sh and tools/gpu_scale.sh on the GPU box): bench_<workload>.log, bench_reference.log,
bench_n<N>_peer.log / bench_n<N>_nccl.log, launches_<workload>.csv (ncu launch lists), prof_*.ncu-rep (ncu --set full).
"""
import collections, csv, io, json, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(R, "gpurun_out"); TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
# on the GPU box the summaries are written next to the raw files (only gpurun_out/ travels back); here into profiles/
P = sys.argv[2] if len(sys.argv) > 2 else os.path.join(R, "profiles")
os.makedirs(P, exist_ok=True)
ONLY = set(sys.argv[3:])       # optional: sections to (re)write -- bench launches full scaling; default all


def section(name):
    return not ONLY or name in ONLY


def bench(name):
    p = os.path.join(G, f"bench_{name}.log")
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def sec_bench():
    lines = [f"# {TAG} -- bench lines on B200 (CUDA-event timings, un-profiled runs)\n"]
    for name in ("default", "poisson", "er", "cari", "rmat"):
        d = bench(name)
        if not d:
            continue
        r = d["roofline"]; wp = r["whole_path"]
        lines.append(f"## {d['config']['workload']}\n")
        lines.append(f"* {d['value']:.1f} GFLOP/s, {d['ms_per_step']:.3f} ms/step (steps {d['steps']}, warmup {d['warmup']}), {d['gpu_launches']} kernel launches in the timed region")
        lines.append(f"* whole path: {wp['algorithmic_bytes']/1e6:.1f} MB algorithmic -> {wp['achieved']:.0f} GB/s = {wp['frac']:.3f} of measured HBM peak ({r['peak']} GB/s)")
        lines.append(f"* dominant compute launch `{r['kernel']}`: {r['kernel_ms']:.3f} ms, {r['algorithmic_bytes']/1e6:.1f} MB algorithmic, {r['achieved']:.0f} GB/s, frac {r['frac']:.3f}; DRAM traffic {('%.0f MB' % (r['traffic']/1e6)) if r.get('traffic') else 'n/a'} ({r.get('timing', 'timed region')})")
        if d.get("setup"):
            lines.append(f"* outside the step: {json.dumps(d['setup'])}")
        if d.get("e2e"):
            lines.append(f"* e2e (host CSR in, C back to host, pinned): {d['e2e']['value']:.2f} GFLOP/s, {d['e2e']['ms_per_step']:.1f} ms/step, H2D {d['e2e']['h2d_bytes_per_step']/1e6:.0f} MB, D2H {d['e2e']['d2h_bytes_per_step']/1e6:.0f} MB")
        if d.get("cpu_baseline"):
            lines.append(f"* CPU oracle (port, {d['cpu_baseline']['cores']} threads): {d['cpu_baseline']['value']:.3f} GFLOP/s on {d['cpu_baseline']['sample']}")
        lines.append(f"* clocks: {d['clocks']}\n")
        lines.append("| launch | ms |\n|---|---|")
        for k, v in r["launch_ms"].items():
            lines.append(f"| {k} | {v:.3f} |")
        lines.append("")
    ref = bench("reference")
    if ref:
        lines.append(f"## reference arm (`bench.py --impl reference`): {ref['value']:.3f} GFLOP/s, {ref['cpu_baseline']}\n")
    open(os.path.join(P, f"{TAG}_bench.md"), "w").write("\n".join(lines))


def sec_launches():
    # launch lists
    PREP = ("k_radix_", "k_transpose_gather", "k_entry_rows", "k_fiber_", "k_validate")   # operand preparation before the steps
    for wl in ("rect", "rmat"):
        ll = os.path.join(G, f"launches_{wl}.csv")
        if not os.path.exists(ll):
            continue
        rows = list(csv.reader(l for l in open(ll) if not l.startswith("==")))
        hdr = rows[0]; ki, vi, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
        agg = collections.OrderedDict()
        for r in rows[1:]:
            if len(r) <= vi:
                continue
            n = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
            if any(x in n for x in PREP):
                continue
            d = agg.setdefault(n, [0, 0.0, r[gi], r[bi]]); d[0] += 1; d[1] += v
        tot = sum(v[1] for v in agg.values())
        out = [f"# {TAG} -- ncu launch list, {wl}, `bench.py --workload {wl} --warmup 3` under ncu\n",
               "`ncu --metrics gpu__time_duration.sum --clock-control none --csv ...` (cold-cache, serialised: compare shares; the operand-preparation kernels that run once before the steps -- fiber store, device transpose, validation -- are left out)\n",
               "| kernel | launches | grid (last) | block | avg us | share |", "|---|---|---|---|---|---|"]
        for k, (n, v, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{k}` | {n} | {g} | {b} | {v/n/1e3:.1f} | {v/tot*100:.1f}% |")
        open(os.path.join(P, f"{TAG}_launches_{wl}.md"), "w").write("\n".join(out) + "\n")


def sec_full():
    # full captures
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
    out = [f"# {TAG} -- `ncu --set full --clock-control none --import-source on` summaries\n",
           "One row per captured launch; times are cold-cache and serialised (profiler replay), the bench lines carry the real ones.\n"]
    traffic = {}
    for rep, wl in (("prof_rect_numeric", "rect"), ("prof_er_fused", "er"), ("prof_poisson_tiny", "poisson"), ("prof_rmat_long", "rmat")):
        p = os.path.join(G, rep + ".ncu-rep")
        if not os.path.exists(p):
            continue
        recs, units = ncu_raw(p)
        out.append(f"## {rep}.ncu-rep ({wl})\n")
        out.append("| kernel | " + " | ".join(w.split(".")[0].replace("__", " ") for w in want) + " |")
        out.append("|---|" + "---|" * len(want))
        for r in recs:
            kn = re.sub(r"\(.*", "", r["Kernel Name"])
            out.append(f"| `{kn}` | " + " | ".join(f"{r.get(w,'')} {units.get(w,'')}".strip() for w in want) + " |")

            def gb(x):
                v = float(r[x].replace(",", "")); return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[x]]
            t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
            traffic.setdefault(wl, {}).setdefault(kn, []).append(t)
        out.append("")
    open(os.path.join(P, f"{TAG}_ncu_full.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(P, f"{TAG}_traffic_raw.json"), "w"), indent=1)


    # traffic.json: DRAM bytes per launch keyed by the engine's launch-record names (what bench.py looks up)
    def record_name(kn):
        m = re.search(r"k_(esc_numeric_warp|bitonic_numeric_cta)<[^,]+, *(\d+)", kn)
        if m:
            return f"sort_pass<{m.group(2)}>"
        m = re.search(r"k_fused_light<[^,]+, *(\d+)", kn)
        if m:
            return f"fused<{m.group(1)}>"
        if "k_fused_tiny" in kn:
            return "fused<32>"
        if "k_long_chunk_sort" in kn:
            return "long_sort"
        if "k_long_reduce" in kn:
            return "long_reduce"
        if "k_copy_rows" in kn:
            return "copy_rows"
        if "k_flops" in kn:
            return "flop_count"
        return None


    named = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the largest captured of that kernel) from the "
                         f"`ncu --set full` captures (profiles/{TAG}_ncu_full.md); keys are the engine's launch-record names; "
                         "records that span several kernels (copy_rows) add their kernels"}
    for wl, ks in traffic.items():
        per = collections.defaultdict(float)
        for kn, ts in ks.items():
            rn = record_name(kn)
            if rn:
                per[rn] += max(ts)
        if per:
            named[wl] = dict(per)
    json.dump(named, open(os.path.join(P, "traffic.json"), "w"), indent=1)


def sec_scaling():
    # scaling table from the torchrun logs
    rows = []
    d1 = bench("default")
    if d1:
        rows.append((1, d1, None))
    for n in (2, 4, 8):
        d = bench(f"n{n}_peer")
        if d:
            rows.append((n, d, bench(f"n{n}_nccl")))
    if len(rows) > 1:
        sc = [f"# {TAG} -- strong scaling of the default workload (rect 1M x 4M, A x A^T), `bench.py --gpus N` under torchrun\n",
              "value = 2 x products / max-over-ranks device time per step with C gathered on every rank by the placement kernels' peer stores;",
              "compute-only = the same steps storing into the rank's own buffers only; NVLink floor = bytes a rank sends / 770 GB/s;",
              "every run checks that each rank's gathered C is bit-identical to the C that rank computes alone.\n",
              "| GPUs | ms/step | GFLOP/s | compute-only ms | NVLink bytes sent per rank | NVLink floor ms | NCCL-gather ms/step (baseline) | C identical | e2e ms | speed-up vs 1 |",
              "|---|---|---|---|---|---|---|---|---|---|"]
        for n, d, dn in rows:
            mg = d.get("multi_gpu") or {}
            sc.append(f"| {n} | {d['ms_per_step']:.3f} | {d['value']:.1f} | {mg.get('compute_only_ms_per_step', d['ms_per_step']):.3f} | "
                      f"{mg.get('nvlink_bytes_sent_per_rank_per_step_max', 0)/1e6:.0f} MB | {mg.get('nvlink_floor_ms', 0):.2f} | "
                      f"{('%.3f' % dn['ms_per_step']) if dn else '-'} | {mg.get('gathered_c_identical_to_one_gpu_on_every_rank', '-')} | "
                      f"{('%.1f' % d['e2e']['ms_per_step']) if d.get('e2e') else '-'} | {rows[0][1]['ms_per_step']/d['ms_per_step']:.2f} |")
        open(os.path.join(P, f"{TAG}_scaling.md"), "w").write("\n".join(sc) + "\n")

for name, fn in (("bench", sec_bench), ("launches", sec_launches), ("full", sec_full), ("scaling", sec_scaling)):
    if section(name):
        fn()
