#!/usr/bin/env python
"""Per-source-line warp-instruction counts of ONE kernel of an ncu report.
usage: python profiles/srcprof2.py rep.ncu-rep <kernel regex> <pairs> [min inst/pair]"""
import csv, subprocess, sys
rep, kre, pairs = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 8.0
out = subprocess.run(["ncu", "-i", rep, "-k", "regex:" + kre, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
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
        rows.append((int(r[0]), r[1], num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")])))
tot, ts = sum(x[2] for x in rows), sum(x[3] for x in rows)
print(f"total warp-instructions {tot}  per pair {tot / pairs:.1f}  stall samples {ts}")
agg = {}
for ln, src, inst, smp in rows:
    a = agg.setdefault((ln, src.strip()[:100]), [0, 0])
    a[0] += inst
    a[1] += smp
for (ln, src) in sorted(agg):
    inst, smp = agg[(ln, src)]
    if inst / pairs >= thr:
        print(f"{ln:4d} {inst / pairs:8.1f} inst/pair {100 * smp / max(ts, 1):5.1f}% samples  {src}")
