#!/usr/bin/env python
"""sass_summary.py LIB: per kernel of the library -- registers, static shared memory, instruction count and the counts of the
mnemonics that show what the code is made of (bulk-copy / mbarrier / shared atomics / SIMD min-max / logic / wide multiplies ...),
from `cuobjdump -sass -res-usage`.  Written to profiles/rN/sass_summary.txt each round."""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1]
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for l in res.splitlines():
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", l)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
        cur = None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(usage), capture_output=True, text=True).stdout.splitlines()
demangled = dict(zip(usage, names))
WATCH = ["UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "ATOMS", "ATOMG", "RED", "VIMNMX", "VIMNMX3", "LOP3", "SHF", "IMAD", "IMAD.WIDE", "POPC", "PRMT",
         "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "SHFL", "VOTE", "REDUX", "MATCH", "BAR", "CCTL", "UTCMMA", "HMMA"]
print(f"# {lib}: cuobjdump -sass -res-usage, sm_100a")
arch = re.search(r"arch = (sm_\w+)", txt)
print("# arch:", arch.group(1) if arch else "?")
for blk in re.split(r"\n\s*Function : ", txt)[1:]:
    name = blk.split("\n", 1)[0].strip()
    ops = Counter()
    n = 0
    for l in blk.splitlines():
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        n += 1
        full = m.group(1)
        base = full.split(".")[0]
        ops[base] += 1
        if full.startswith("IMAD.WIDE"):
            ops["IMAD.WIDE"] += 1
    reg, sh = usage.get(name, (None, None))
    d = demangled.get(name, name)
    cut = d.rfind(">(")
    d = d[:cut + 1] if cut > 0 else d.split("(")[0]
    d = d.replace("(int)", "").replace("(bool)", "")
    print(f"\n{d}\n  registers {reg}  static smem {sh} B  SASS instructions {n}")
    print("  " + "  ".join(f"{k} {ops[k]}" for k in WATCH if ops[k]))
