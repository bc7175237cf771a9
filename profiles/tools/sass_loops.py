#!/usr/bin/env python
"""sass_loops.py LIB KERNEL_SUBSTRING: loops of a kernel (backward branches) with their instruction mix, from cuobjdump -sass."""
import re
import subprocess
import sys
from collections import Counter

lib, want = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0].strip()
    if want not in name:
        continue
    print("==", name)
    ins, labels, cur = [], {}, None
    for l in blk.split("\n"):
        m = re.match(r"\s*(\.L_x_\d+):", l)
        if m:
            cur = m.group(1)
            continue
        m2 = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m2:
            a = int(m2.group(1), 16)
            ins.append((a, m2.group(2).strip()))
            if cur:
                labels[cur] = a
                cur = None
    print("instructions:", len(ins))
    def op(t):
        t = t.split()
        t = t[1] if t[0].startswith("@") else t[0]
        return t.split(".")[0]
    print("total mix:", dict(Counter(op(t) for _, t in ins).most_common(14)))
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:`\((\.L_x_\d+)\)|(0x[0-9a-f]+))", t)
        if not m:
            continue
        tgt = labels.get(m.group(1)) if m.group(1) else int(m.group(2), 16)
        if tgt is not None and tgt < a:
            body = [x for x in ins if tgt <= x[0] <= a]
            print(f"loop {tgt:#x} -> {a:#x}: {len(body)} instructions", dict(Counter(op(x[1]) for x in body).most_common(12)))
