#!/usr/bin/env python
"""Aggregate an ncu report per code region (function / '---- ' section) of pb_kernels.cuh.
usage: python profiles/regionprof.py rep.ncu-rep <pairs> [source file]"""
import collections, csv, subprocess, sys
rep, pairs = sys.argv[1], float(sys.argv[2])
srcf = sys.argv[3] if len(sys.argv) > 3 else "pandaseq_b200/csrc/pb_kernels.cuh"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
hdr, rows = None, []
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        rows.append((int(r[0]), num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")])))
src = open(srcf).read().split("\n")
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(src)
         if l.startswith("__device__") or l.startswith("template") or l.startswith("__global__") or "/* ----" in l]
def region(ln):
    name = "?"
    for m, l in marks:
        if m <= ln:
            name = f"{m}:{l}"
    return name
agg, ts, ti = collections.OrderedDict(), sum(r[2] for r in rows), sum(r[1] for r in rows)
for ln, inst, smp in sorted(rows):
    a = agg.setdefault(region(ln), [0, 0])
    a[0] += inst
    a[1] += smp
print(f"total warp-instructions per pair {ti / pairs:.1f}")
for k, (i, s) in agg.items():
    if i / pairs >= 3:
        print(f"{i / pairs:8.1f} inst/pair {100 * s / max(ts, 1):5.1f}% samples  {k}")
