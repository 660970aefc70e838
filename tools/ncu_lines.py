#!/usr/bin/env python
"""Summarise an .ncu-rep per CUDA source line: warp instructions executed and stall samples.
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
items = []; fname = "?"; hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0] or not r[0].isdigit(): continue
    ie_i = hdr.index("Instructions Executed"); sm_i = hdr.index("# Samples")
    try: ie = int(float(r[ie_i])); sm = int(float(r[sm_i]))
    except ValueError: continue
    if ie or sm: items.append((ie, sm, fname, int(r[0]), r[1].strip()[:100]))
tot = sum(i[0] for i in items) or 1; ts = sum(i[1] for i in items) or 1
print(f"total warp instructions {tot:,}   samples {ts:,}")
for ie, sm, f, ln, src in sorted(items, reverse=True)[:top]:
    print(f"{ie/tot*100:5.1f}% inst {sm/ts*100:5.1f}% smpl  {f}:{ln}: {src}")
