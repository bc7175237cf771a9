#!/usr/bin/env python
"""dram_traffic.py PREFIX OUT.json: per BASELINE config, the DRAM bytes per pair of the kernels of one step, from the ncu csv files
PREFIX_dram_cfgN.csv (profiles/tools/dram_traffic.sh).  The first step's launches are taken: from the first kernel of the path up to
(not including) its second occurrence."""
import csv
import json
import re
import sys

prefix, out = sys.argv[1], sys.argv[2]
PAIRS = 2_000_000
res = {}
for c in (2, 3, 4, 5):
    try:
        lines = [l for l in open(f"{prefix}_dram_cfg{c}.csv") if l.startswith('"')]
    except OSError:
        continue
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        k = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
        k[r["Metric Name"]] = v * scale
    ids = sorted(launches)
    if not ids:
        continue
    def short(n):
        m = re.search(r"(pb[a-z]*::)?([a-z_]+kernel)(<[^>]*>)?", n)
        return (m.group(2) + (m.group(3) or "")) if m else n[:40]
    first = short(launches[ids[0]]["name"])
    step = [ids[0]]
    for i in ids[1:]:
        if short(launches[i]["name"]) == first:
            break
        step.append(i)
    per = {}
    for i in step:
        l = launches[i]
        d = per.setdefault(short(l["name"]), {"read": 0.0, "write": 0.0, "us": 0.0, "launches": 0})
        d["read"] += l.get("dram__bytes_read.sum", 0.0) / PAIRS
        d["write"] += l.get("dram__bytes_write.sum", 0.0) / PAIRS
        d["us"] += l.get("gpu__time_duration.sum", 0.0)
        d["launches"] += 1
    rd = sum(d["read"] for d in per.values())
    wr = sum(d["write"] for d in per.values())
    for d in per.values():
        d["read"], d["write"], d["us"] = round(d["read"], 1), round(d["write"], 1), round(d["us"], 1)
    res[str(c)] = {"dram_bytes_per_pair": round(rd + wr, 1), "read": round(rd, 1), "write": round(wr, 1), "pairs": PAIRS, "per_kernel": per,
                   "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on bench.py --config {c} --pairs {PAIRS} "
                             "(profiles/tools/dram_traffic.sh): the kernels of the first step; durations are ncu's (cold cache, serialised)"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: (v["read"], v["write"]) for k, v in res.items()}))
