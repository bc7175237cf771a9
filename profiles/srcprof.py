#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: `python profiles/srcprof.py rep.ncu-rep <pairs> [min_inst_per_pair]`."""
import csv, subprocess, sys
rep, pairs = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
hdr, rows = None, []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        def num(x):
            try:
                return int(x)
            except ValueError:
                return 0
        rows.append((int(r[0]), r[1], num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")])))
tot, ts = sum(x[2] for x in rows), sum(x[3] for x in rows)
print(f"total warp-instructions {tot}  per pair {tot / pairs:.1f}  stall samples {ts}")
agg = {}
for ln, src, inst, smp in rows:
    a = agg.setdefault(ln, [src, 0, 0])
    a[1] += inst
    a[2] += smp
for ln in sorted(agg):
    src, inst, smp = agg[ln]
    if inst / pairs >= thr:
        print(f"{ln:4d} {inst / pairs:8.1f} inst/pair {100 * smp / max(ts, 1):5.1f}% samples  {src.strip()[:110]}")
