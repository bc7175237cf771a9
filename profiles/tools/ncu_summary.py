#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep [pairs]: the metrics this project reads from an `ncu --set full` capture, one block per kernel launch."""
import csv
import subprocess
import sys

rep = sys.argv[1]
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("=====", r[idx["Kernel Name"]][:110])
    for w in want:
        if w in idx:
            print(f"  {w} = {r[idx[w]]} {units[idx[w]]}")
    if pairs:
        inst = float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        print(f"  warp-instructions per pair = {inst / pairs:.1f}; DRAM bytes per pair = {rd * scale[units[idx['dram__bytes_read.sum']]] / pairs:.0f} read + "
              f"{wr * scale[units[idx['dram__bytes_write.sum']]] / pairs:.0f} written")
    st = [(float(r[i].replace(",", "")), h) for h, i in idx.items()
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i]]
    print("  stalls (warps per issue): " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in sorted(st, reverse=True)[:6]))
