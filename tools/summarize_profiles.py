#!/usr/bin/env python
"""Turns gpurun_out/ (scratch) into the tracked summaries under profiles/ for this round."""
import collections, csv, io, json, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(R, "gpurun_out"); P = os.path.join(R, "profiles"); TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

def bench(name):
    p = os.path.join(G, f"bench_{name}.log")
    return json.loads(open(p).read().strip().splitlines()[-1]) if os.path.exists(p) else None

def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))

lines = [f"# {TAG} -- bench lines on B200 (CUDA-event timings, un-profiled runs)\n"]
traffic = {}
for name in ("default", "poisson", "er", "cari", "rmat"):
    d = bench(name)
    if not d: continue
    r = d["roofline"]; wp = r["whole_path"]
    lines.append(f"## {d['config']['workload']}\n")
    lines.append(f"* {d['value']:.1f} GFLOP/s, {d['ms_per_step']:.3f} ms/step (steps {d['steps']}, warmup {d['warmup']}), {d['gpu_launches']} kernel launches in the timed region")
    lines.append(f"* whole path: {wp['algorithmic_bytes']/1e6:.1f} MB algorithmic -> {wp['achieved']:.0f} GB/s = {wp['frac']:.3f} of measured HBM peak ({r['peak']} GB/s)")
    lines.append(f"* dominant launch `{r['kernel']}`: {r['kernel_ms']:.3f} ms, {r['achieved']:.0f} GB/s, frac {r['frac']:.3f} ({r.get('timing', 'timed region')})")
    if d["config"].get("device_transpose_ms"): lines.append(f"* device transpose of A (B = A^T, `spada_b200_transpose`): {d['config']['device_transpose_ms']:.2f} ms")
    if d.get("e2e"): lines.append(f"* e2e (host CSR in, C back to host, pinned): {d['e2e']['value']:.2f} GFLOP/s, {d['e2e']['ms_per_step']:.1f} ms/step, H2D {d['e2e']['h2d_bytes_per_step']/1e6:.0f} MB, D2H {d['e2e']['d2h_bytes_per_step']/1e6:.0f} MB")
    if d.get("cpu_baseline"): lines.append(f"* CPU oracle (port, {d['cpu_baseline']['cores']} threads): {d['cpu_baseline']['value']:.3f} GFLOP/s on {d['cpu_baseline']['sample']}")
    lines.append(f"* clocks: {d['clocks']}\n")
    lines.append("| launch | ms |\n|---|---|")
    for k, v in r["launch_ms"].items(): lines.append(f"| {k} | {v:.3f} |")
    lines.append("")
ref = bench("reference")
if ref: lines.append(f"## reference arm (`bench.py --impl reference`): {ref['value']:.3f} GFLOP/s, {ref['cpu_baseline']}\n")
open(os.path.join(P, f"{TAG}_bench.md"), "w").write("\n".join(lines))

# launch list
ll = os.path.join(G, "launches_rect.csv")
if os.path.exists(ll):
    rows = list(csv.reader(l for l in open(ll) if not l.startswith("==")))
    hdr = rows[0]; ki, vi, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    PREP = ("k_radix_", "k_transpose_gather", "k_entry_rows", "k_fiber_", "k_validate")   # operand preparation before the steps
    for r in rows[1:]:
        if len(r) <= vi: continue
        n = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
        if any(x in n for x in PREP): continue
        d = agg.setdefault(n, [0, 0.0, r[gi], r[bi]]); d[0] += 1; d[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# {TAG} -- ncu launch list, default workload (rect), `bench.py --workload rect --steps 2 --warmup 3`\n",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv ...` (cold-cache, serialised: compare shares; the operand-preparation kernels that run once before the steps -- fiber store, device transpose -- are left out)\n",
           "| kernel | launches | grid | block | avg us | share |", "|---|---|---|---|---|---|"]
    for k, (n, v, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {g} | {b} | {v/n/1e3:.1f} | {v/tot*100:.1f}% |")
    open(os.path.join(P, f"{TAG}_launches_rect.md"), "w").write("\n".join(out) + "\n")

# full captures
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
out = [f"# {TAG} -- `ncu --set full --clock-control none --import-source on` summaries\n"]
NOTES = {"prof_rmat_huge": "captured BEFORE the sub-batch change of heavy.cu (`item_sub_batch`): the evidence for it -- "
                           "one warp of eight busy per item, 9-15 % active warps, 3 % issue slots"}
for rep, wl in (("prof_rect_heavy", "rect"), ("prof_rect_numeric", "rect"), ("prof_er_fused", "er"), ("prof_poisson_tiny", "poisson"),
                ("prof_rmat_huge", "rmat_before_sub_batch")):  # noqa
    p = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(p): continue
    recs, units = ncu_raw(p)
    out.append(f"## {rep}.ncu-rep ({wl})\n")
    if rep in NOTES: out.append(NOTES[rep] + "\n")
    out.append("| kernel | " + " | ".join(w.split(".")[0].replace("__", " ") for w in want) + " |")
    out.append("|---|" + "---|" * len(want))
    for r in recs:
        kn = re.sub(r"\(.*", "", r["Kernel Name"])
        out.append(f"| `{kn}` | " + " | ".join(f"{r.get(w,'')} {units.get(w,'')}".strip() for w in want) + " |")
        def gb(x, u): 
            v = float(r[x].replace(",", "")); return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        t = gb("dram__bytes_read.sum", units["dram__bytes_read.sum"]) + gb("dram__bytes_write.sum", units["dram__bytes_write.sum"])
        traffic.setdefault(wl, {})[kn] = t
    out.append("")
open(os.path.join(P, f"{TAG}_ncu_full.md"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(P, f"{TAG}_traffic_raw.json"), "w"), indent=1)

# traffic.json: the same numbers keyed by the engine's launch-record names (what bench.py looks up)
def record_name(kn):
    m = re.search(r"k_(esc_numeric_warp|bitonic_numeric_cta)<[^,]+, *(\d+)", kn)
    if m: return f"sort_pass<{m.group(2)}>"
    m = re.search(r"k_fused_light<[^,]+, *(\d+)", kn)
    if m: return f"fused<{m.group(1)}>"
    if "k_fused_tiny" in kn: return "fused<32>"
    if "k_heavy_smem_numeric" in kn: return "oneshot<heavy>"
    if "k_copy_rows" in kn: return "copy_rows"
    return None
named = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures "
                     f"(profiles/{TAG}_ncu_full.md); keys are the engine's launch-record names; copy_rows = both copy kernels"}
for wl, ks in traffic.items():
    for kn, t in ks.items():
        rn = record_name(kn)
        if rn: named.setdefault(wl, {})[rn] = named.get(wl, {}).get(rn, 0) + t
json.dump(named, open(os.path.join(P, "traffic.json"), "w"), indent=1)

# scaling table from the torchrun logs
rows = []
d1 = bench("default")
if d1: rows.append((1, d1["ms_per_step"], d1["value"], d1["compute_only"]["ms_per_step"]))
for n in (2, 4, 8):
    pth = os.path.join(G, f"multi_rect_{n}_w1.log")
    if os.path.exists(pth):
        l = [x for x in open(pth) if x.startswith("{")]
        if l:
            d = json.loads(l[-1]); rows.append((n, d["ms_per_step"], d["value"], d["compute_only"]["ms_per_step"]))
if len(rows) > 1:
    sc = [f"# {TAG} -- strong scaling of the default workload (rect 1M x 4M, A x A^T), `bench.py --gpus N` under torchrun\n",
          "value = 2 x products / max-over-ranks device time per step, C all-gathered on every rank; compute-only = the engine's own",
          "device time (max over ranks) without the gather.\n", "| GPUs | ms/step | GFLOP/s | compute-only ms | speed-up vs 1 |", "|---|---|---|---|---|"]
    for n, ms, v, co in rows: sc.append(f"| {n} | {ms:.3f} | {v:.1f} | {co:.3f} | {rows[0][1]/ms:.2f} |")
    open(os.path.join(P, f"{TAG}_scaling.md"), "w").write("\n".join(sc) + "\n")
print(open(os.path.join(P, f"{TAG}_bench.md")).read()[:3000])
