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
for name in ("default", "poisson", "er", "cari"):
    d = bench(name)
    if not d: continue
    r = d["roofline"]; wp = r["whole_path"]
    lines.append(f"## {d['config']['workload']}\n")
    lines.append(f"* {d['value']:.1f} GFLOP/s, {d['ms_per_step']:.3f} ms/step (steps {d['steps']}, warmup {d['warmup']}), {d['gpu_launches']} kernel launches in the timed region")
    lines.append(f"* whole path: {wp['algorithmic_bytes']/1e6:.1f} MB algorithmic -> {wp['achieved']:.0f} GB/s = {wp['frac']:.3f} of measured HBM peak ({r['peak']} GB/s)")
    lines.append(f"* dominant launch `{r['kernel']}`: {r['kernel_ms']:.3f} ms, {r['achieved']:.0f} GB/s, frac {r['frac']:.3f}")
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
    for r in rows[1:]:
        if len(r) <= vi: continue
        n = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
        d = agg.setdefault(n, [0, 0.0, r[gi], r[bi]]); d[0] += 1; d[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# {TAG} -- ncu launch list, default workload (rect), `bench.py --workload rect --steps 2 --warmup 3`\n",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv ...` (cold-cache, serialised: compare shares)\n",
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
for rep, wl in (("prof_rect_numeric", "rect"), ("prof_er_fused", "er"), ("prof_poisson_tiny", "poisson")):  # noqa
    p = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(p): continue
    recs, units = ncu_raw(p)
    out.append(f"## {rep}.ncu-rep ({wl})\n")
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
print(open(os.path.join(P, f"{TAG}_bench.md")).read()[:3000])
