#!/usr/bin/env python
"""bench.py -- read-pairs/s of the paired-end assembly hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2] [--pairs P] [--path pairs|text|api|copy]

A "step" is one pass of the hot path (primer scan when configured, k-mer seeded overlap selection,
reconstruction with posterior qualities) over one batch of synthetic read pairs.  At N = 1 the batch is
BASELINE config 2: 10 M synthetic 2x150 bp pairs, simple_bayesian.  With N > 1 (launched by torchrun,
one rank per GPU) every rank assembles its own 10 M-pair shard -- weak scaling, no collective on the
data path; the per-rank counter vectors are merged once at the end (the STAT merge).

value      whole-job Mpairs/s with the packed batch already resident in HBM (CUDA events on the
           library's launch stream, max over ranks).
e2e        the same metric through the C ABI with HOST buffers: pb_assemble_host() copies the flat
           panda_qual arrays host->device, packs, assembles, and copies results + merged reads back.
roofline   algorithmic HBM read bytes per pair (458 B at 2x150) x pairs / summed duration of the step's kernels vs
           the measured copy bandwidth in MEASURED_PEAKS.json; `kernels` lists what a step launches (the seeding
           kernel, the bin list, the lane-per-pair kernel and the general kernel's list pass for the common
           configurations; the general kernel alone otherwise) with each one's duration from the library's own
           CUDA events (pb_set_timing / pb_last_timing) and its own algorithmic bytes.
cpu_baseline / --impl reference
           the reference's own CPU implementation (oracle/_ref, compiled from the reference sources)
           or the oracle port if that is absent, all host cores, on a bounded sample of the same workload.
--path api the same metric through the reference's OWN entry points -- a PandaNextSeq source, panda_run_pool(), a
           PandaOutputSeq callback that reads every base of every result (tools/api_bench.c, one source file built against
           this library and against the compiled reference) -- for 1 .. all host threads on both sides.
--path copy
           what the platform gives the host path: pinned host <-> device copies of the e2e leg's sizes, both directions at
           once, no kernels, on every rank at the same time (the ceiling e2e is measured against).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GEN_CHUNK = 1_000_000


def algorithmic_read_bytes(flen, rlen):
    """4-bit nt + 8-bit PHRED for both reads + 8 B of pair metadata."""
    return (flen + 1) // 2 + (rlen + 1) // 2 + flen + rlen + 8


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms while the kernel under test is running."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(local):
    """Run this rank on the host cores next to its GPU (sysfs: the PCI device's local_cpulist / numa_node), so that the page-locked
    buffers it allocates afterwards land in that socket's memory and the copies do not cross the socket link.  Done before the
    CUDA context and every pinned allocation; silently a no-op where the container's cpuset leaves no choice."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        dev = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{dev}"
        info["numa_node"] = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        near = cpus & allowed
        info["cpus"] = len(near)
        info["allowed_cpus"] = len(allowed)
        if near and near != allowed:
            os.sched_setaffinity(0, near)
            info["bound"] = True
    except Exception as e:          # no sysfs entry, no permission: run unbound
        info["error"] = str(e)[:80]
    return info


def workload(cfg_id):
    from pandaseq_b200 import synth
    c = synth.CONFIGS[cfg_id]
    kw = {}
    if c.get("primers"):
        fwd = synth.encode(synth.FWD_PRIMER)
        rev = synth.encode("".join(synth._COMP[ch] for ch in synth.REV_PRIMER))
        kw = dict(forward_primer=fwd, reverse_primer=rev)
    return c, kw


def shape_label(c):
    """read lengths of a BASELINE config as the config states them (not as pair 0 happens to have them)"""
    lo, hi = c["rl"]
    return f"2x({lo}-{hi}) bp mixed lengths" if c.get("mixed") else f"2x{lo} bp"


def pairs_per_gpu(c, args):
    """BASELINE sizes: 10 M pairs per GPU; config 5 is 100 M pairs over 8 GPUs, i.e. 12.5 M per GPU (weak scaling: the same at every N)"""
    if args.pairs is not None:
        return args.pairs
    return 12_500_000 if c.get("mixed") else min(c["n"], 10_000_000)


def cpu_rate(cfg, flat, threads):
    """Mpairs/s of the CPU reference on `flat`, and which implementation ran."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    which = "ref" if oracle_lib.have_ref() else "port"
    t0 = time.perf_counter()
    oracle_lib.assemble(which, cfg, flat, want_seq=False, threads=threads)
    dt = time.perf_counter() - t0
    return flat.n / dt / 1e6, ("reference" if which == "ref" else "port")


def reference_arm(args):
    """--impl reference: the reference's own CPU path on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pandaseq_b200 as pb
    from pandaseq_b200 import synth
    c, kw = workload(args.config)
    cfg = pb.make_config(c["algo"], **kw)
    cores = os.cpu_count() or 1
    probe = synth.generate_config(args.config, n=20_000, device="cpu").to_flat()
    rate, kind = cpu_rate(cfg, probe, cores)
    sample = int(min(max(rate * 1e6 * 6.0, 20_000), 10_000_000 * 300 // (2 * c["rl"][1])))        # ~6 s per step, at most the config's 10 M (fewer of longer reads: host memory)
    flat = synth.generate_config(args.config, n=sample, device="cpu", chunk_index=1).to_flat()
    for _ in range(args.warmup):
        cpu_rate(cfg, flat.slice(0, min(sample, 50_000)), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_rate(cfg, flat, cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt / 1e6
    fl, rl = flat.lengths()
    line = {
        "impl": "reference", "metric": "read-pairs/s (Mpairs/s), pair assembly hot path", "value": value, "unit": "Mpairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config {args.config}: synthetic {shape_label(c)} pairs, {c['algo']}"
                               + (", primer strip" if kw else "") + f", bounded sample of {sample} pairs per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": "Mpairs/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} pairs x {args.steps} steps, panda_assembler_assemble loop, one assembler per thread, logging off"},
        "e2e": {"value": value, "unit": "Mpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def text_path(args):
    """--path text: the chain FASTQ text -> parse -> assemble -> FASTA text (pb_fastq_assemble_host, pinned host buffers),
    its three device stages timed separately on resident text, and the reference's own CLI on a bounded sample."""
    import torch
    import pandaseq_b200 as pb
    from pandaseq_b200 import synth
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    c, kw = workload(args.config)
    cfg = pb.make_config(c["algo"], **kw)
    n = args.pairs if args.pairs is not None else 4_000_000
    ctx = pb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=dev)
    ftxt, rtxt = [], []
    for ci, start in enumerate(range(0, n, GEN_CHUNK)):
        cnt = min(GEN_CHUNK, n - start)
        rect = synth.generate_config(args.config, n=cnt, device=dev, chunk_index=ci)
        f_data, f_off, r_data, r_off = rect.to_flat_tensors()
        ftxt.append(synth.fastq_text(f_data, f_off, 1, complement=False, seed=7 + ci))
        rtxt.append(synth.fastq_text(r_data, r_off, 2, complement=True, seed=7 + ci))
        del rect, f_data, r_data
    # ---- device-resident stages on the first 1 M pairs (one chunk of the chain) -------------------------
    df, dr = ftxt[0], rtxt[0]
    n1 = min(n, GEN_CHUNK)
    import ctypes as C
    stage_ms = {"parse": [], "assemble": [], "format": []}
    parsed = ctx.fastq_parse_device(df, dr, max_records=n1 + 16)
    info = parsed["info"]
    lim = int(info["limit"])
    stride = (2 * int(info["max_read_len"]) + 15) & ~15
    results = torch.empty((lim, 32), dtype=torch.uint8, device=dev)
    nt = torch.empty((lim, stride // 2), dtype=torch.uint8, device=dev)
    counters = torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64, device=dev)
    text_dev = torch.empty(int(df.numel()), dtype=torch.uint8, device=dev)
    total = C.c_size_t(0)
    L = pb.lib()
    for it in range(args.warmup + args.steps):
        # parse and format return after a small device->host read, so the host clock around the call is the stage time
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        parsed = ctx.fastq_parse_device(df, dr, max_records=n1 + 16)
        t_parse = (time.perf_counter() - t0) * 1e3
        e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e1.record(stream)
        ctx.assemble_device(cfg, lim, int(info["max_read_len"]), parsed["reads"], parsed["meta"], results, nt, None, stride, counters)
        e2.record(stream)
        ctx.synchronize()
        t0 = time.perf_counter()
        rc = L.pb_format_device(ctx._h, pb.OUT_FASTA, lim, results.data_ptr(), nt.data_ptr(), None, stride, parsed["ids"].data_ptr(),
                                df.data_ptr(), text_dev.data_ptr(), text_dev.numel(), C.byref(total))
        t_fmt = (time.perf_counter() - t0) * 1e3
        if rc != 0:
            raise RuntimeError(L.pb_last_error().decode())
        if it >= args.warmup:
            stage_ms["parse"].append(t_parse)
            stage_ms["assemble"].append(e1.elapsed_time(e2))
            stage_ms["format"].append(t_fmt)
    del results, nt, text_dev
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    text_bytes = int(df.numel() + dr.numel())
    rec_bytes = int(parsed["info"]["stride16"]) * 16 * n1
    parse_alg = text_bytes + rec_bytes + (48 + 8) * n1
    parse_ms = float(np.median(stage_ms["parse"]))
    # ---- the chain with pinned host buffers ---------------------------------------------------------------------
    hf = torch.cat(ftxt).cpu().pin_memory()
    hr = torch.cat(rtxt).cpu().pin_memory()
    del ftxt, rtxt, df, dr
    out = torch.zeros(hf.numel() + hr.numel(), dtype=torch.uint8).pin_memory()
    infos = None
    for _ in range(max(1, min(args.warmup, 2))):
        _, infos, cnt = ctx.fastq_assemble_host(cfg, hf, hr, out=out)
    torch.cuda.synchronize(dev)
    ksteps = max(1, min(args.steps, 3))
    sampler = ClockSampler(local)
    t0 = time.perf_counter()
    for _ in range(ksteps):
        _, infos, cnt = ctx.fastq_assemble_host(cfg, hf, hr, out=out)
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    assert infos["pairs"] == n and infos["error"] == 0, infos
    value = n * ksteps / dt / 1e6
    # ---- the reference's CLI on a bounded sample of the same text -------------------------------------------------
    cpu = None
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "pandaseq")
    if not args.no_cpu and os.path.exists(ref_bin):
        import tempfile
        sample = min(n, 400_000)
        fb = bytes(hf.numpy()).split(b"\n", 4 * sample)
        rb = bytes(hr.numpy()).split(b"\n", 4 * sample)
        with tempfile.TemporaryDirectory() as td:
            pf_, pr_ = os.path.join(td, "f.fastq"), os.path.join(td, "r.fastq")
            open(pf_, "wb").write(b"\n".join(fb[:4 * sample]) + b"\n")
            open(pr_, "wb").write(b"\n".join(rb[:4 * sample]) + b"\n")
            cores = os.cpu_count() or 1
            best = None
            for threads in sorted({1, min(cores, 8), min(cores, 32)}):
                t0 = time.perf_counter()
                subprocess.run([ref_bin, "-f", pf_, "-r", pr_, "-A", c["algo"], "-T", str(threads), "-w", os.devnull, "-g", os.devnull],
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                rate = sample / (time.perf_counter() - t0) / 1e6
                if best is None or rate > best[0]:
                    best = (rate, threads)
            cpu = {"value": best[0], "unit": "Mpairs/s", "cores": best[1], "kind": "reference",
                   "sample": f"the reference's own CLI (oracle/_ref/pandaseq -f -r -T {best[1]} -w /dev/null) on the first {sample} pairs of the same "
                             "FASTQ text; best of -T 1/8/32 (it is parser-bound behind a mutex and does not scale)"}
    line = {
        "metric": "read-pairs/s (Mpairs/s), FASTQ text in -> assembled FASTA text out", "value": value, "unit": "Mpairs/s", "n_gpus": 1,
        "steps": ksteps, "warmup": min(args.warmup, 2), "ms_per_step": dt / ksteps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 text -> u8/f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config {args.config} as FASTQ text: {n} pairs, {c['algo']}, CASAVA 1.7 headers, PHRED+33; "
                               "pinned host text in, pinned host FASTA text out (pb_fastq_assemble_host)",
                   "pairs": n, "l2_policy": "inputs larger than L2, streamed in 96 MB windows"},
        "e2e": {"value": value, "unit": "Mpairs/s", "h2d_bytes_per_step": int(hf.numel() + hr.numel()), "d2h_bytes_per_step": int(infos["out_bytes"])},
        "stages_resident": {"pairs": n1, "parse_ms": parse_ms, "assemble_ms": float(np.median(stage_ms["assemble"])),
                            "format_ms": float(np.median(stage_ms["format"])),
                            "note": "one 1 M-pair chunk with its text resident in HBM; parse = line index + identifiers + reads + finish "
                                    "(host wall clock around pb_fastq_parse_device, which returns after its 64-byte info came back)"},
        "roofline": {"bound": "hbm", "kernel": "pbio::nl_count+nl_write+fq_ids+fq_reads (FASTQ parse)", "achieved": parse_alg / (parse_ms / 1e3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": parse_alg / (parse_ms / 1e3) / 1e9 / peak, "traffic": None,
                     "algorithmic_bytes": parse_alg, "note": "text read once + packed records, identifiers and metadata written once"},
        "cpu_baseline": cpu, "gpu_launches": None, "clocks": clocks,
        "stat": {"count": int(cnt[pb.C_COUNT]), "ok": int(cnt[pb.C_OK]), "out_bytes": int(infos["out_bytes"])},
    }
    print(json.dumps(line), flush=True)
    ctx.close()


def api_path(args):
    """--path api: pairs per second through panda_run_pool() with a PandaNextSeq source and a PandaOutputSeq callback that reads
    every base of every result -- tools/api_bench.c, the same source built against this library (pandaseq_b200/api_bench) and
    against the compiled reference (oracle/_ref/api_bench_ref), for 1 .. all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ours, ref = os.path.join(ROOT, "pandaseq_b200", "api_bench"), os.path.join(ROOT, "oracle", "_ref", "api_bench_ref")
    cores = os.cpu_count() or 1
    n = args.pairs if args.pairs is not None else 4_000_000
    rl = 150
    rows = {"ours": [], "reference": []}
    tlist = sorted({1, 2, 4, 8, min(16, cores), min(32, cores)})
    for t in tlist:
        out = subprocess.run([ours, str(n), str(rl), str(t)], check=True, capture_output=True, text=True).stdout
        rows["ours"].append(json.loads(out.strip().split("\n")[-1]))
    if os.path.exists(ref) and not args.no_cpu:
        for t in tlist:
            nr = min(n, max(200_000, 150_000 * t))          # ~ 2-6 s of CPU work per run
            out = subprocess.run([ref, str(nr), str(rl), str(t)], check=True, capture_output=True, text=True).stdout
            rows["reference"].append(json.loads(out.strip().split("\n")[-1]))
    best = max(rows["ours"], key=lambda r: r["mpairs_per_s"])
    best_ref = max(rows["reference"], key=lambda r: r["mpairs_per_s"]) if rows["reference"] else None
    # the two libraries were given the same generator: the number of accepted pairs per input pair must agree
    same = None
    if best_ref:
        a, b = rows["ours"][0], rows["reference"][0]
        same = abs(a["ok"] / a["pairs"] - b["ok"] / b["pairs"]) < 1e-3
    line = {
        "metric": "read-pairs/s (Mpairs/s) through panda_run_pool / panda_assembler_next with host callbacks", "value": best["mpairs_per_s"],
        "unit": "Mpairs/s", "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 + f64", "data": "synthetic",
        "config": {"workload": f"{n} synthetic 2x{rl} bp pairs, simple_bayesian, pulled one pair at a time from a PandaNextSeq source; every base "
                               "and log p of every result read by the PandaOutputSeq callback (tools/api_bench.c)", "threads": best["threads"]},
        "e2e": {"value": best["mpairs_per_s"], "unit": "Mpairs/s", "h2d_bytes_per_step": int(n * (4 * rl + 16 + 4)),
                "d2h_bytes_per_step": int(n * (32 + (2 * rl + 15) // 16 * 16 // 2 + (2 * rl + 15) // 16 * 16 * 2))},
        "by_threads": rows, "same_outcome_as_reference": same,
        "cpu_baseline": None if not best_ref else {"value": best_ref["mpairs_per_s"], "unit": "Mpairs/s", "cores": best_ref["threads"], "kind": "reference",
                                                   "sample": "the reference's own panda_run_pool + PandaMux over the same source and callback "
                                                             f"(oracle/_ref/api_bench_ref), best of {tlist} threads"},
        "gpu_launches": None,
    }
    print(json.dumps(line), flush=True)


def copy_path(args):
    """--path copy: pinned host -> device and device -> host copies of the sizes the e2e leg moves per chunk, both directions at
    once on two streams, no kernels; every rank at the same time (max over ranks).  The platform's ceiling for the host path."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_near_gpu(local) if os.environ.get("PANDASEQ_B200_NOBIND") is None else {"bound": False}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    pairs = 1 << 18                                   # one chunk of the host path
    h2d_b, d2h_b = pairs * 620, pairs * 184           # AoS in, result record + merged read out (2x150)
    hin = torch.empty(h2d_b, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_b, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_b, dtype=torch.uint8, device=dev)
    dout = torch.empty(d2h_b, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        chunks = 16 * max(1, args.steps)
        for it in range(2):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(chunks):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        din.copy_(hin, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        res[mode] = {"seconds": dt, "h2d_gbs_per_gpu": (h2d_b * chunks / dt / 1e9) if mode != "d2h" else 0.0,
                     "d2h_gbs_per_gpu": (d2h_b * chunks / dt / 1e9) if mode != "h2d" else 0.0,
                     "mpairs_per_s_all_gpus": pairs * chunks * world / dt / 1e6}
    if rank == 0:
        line = {"metric": "pinned host <-> device copies of the host path's chunk sizes (no kernels)", "value": res["both"]["mpairs_per_s_all_gpus"],
                "unit": "Mpairs/s equivalent", "n_gpus": world, "steps": args.steps, "warmup": 1, "higher_is_better": True, "scaling": "weak",
                "config": {"workload": f"{pairs} pairs per chunk: {h2d_b} B in (620 B per pair), {d2h_b} B out (184 B per pair), 2 streams per GPU"},
                "modes": res, "host_affinity_rank0": affinity}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU (default: the config's, capped at 10 M per GPU)")
    ap.add_argument("--path", default="pairs", choices=["pairs", "text", "api", "copy"],
                    help="pairs: the BASELINE metric on panda_qual pairs (default); text: FASTQ text in -> FASTA text out (SURVEY.md 8f rank 1+2); "
                         "api: through panda_run_pool and the reference's callbacks; copy: host <-> device copies alone")
    ap.add_argument("--per-base-p", action="store_true",
                    help="also produce the per-base log p of every merged base (what FASTQ output and the quality plugins need): doubles on the "
                         "device-resident leg, 16-bit codes into the posterior table on the e2e leg (pb_assemble_host_codes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.path == "text":
        text_path(args)
        return
    if args.path == "api":
        api_path(args)
        return
    if args.path == "copy":
        copy_path(args)
        return

    import torch
    import pandaseq_b200 as pb
    from pandaseq_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    affinity = bind_near_gpu(local) if os.environ.get("PANDASEQ_B200_NOBIND") is None else {"bound": False}
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    c, kw = workload(args.config)
    cfg = pb.make_config(c["algo"], **kw)
    n = pairs_per_gpu(c, args)
    ctx = pb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=dev)

    # ---- build the resident batch: generate on the device, pack with the library's pack kernel ------
    reads_parts, meta_parts, alg_bytes, max_len, base16 = [], [], 0, 0, 0
    keep_flat = []          # AoS chunks kept (on the host) for the e2e leg
    # host memory of the e2e and CPU legs is budgeted in bases, not pairs: 4 M pairs of 2x150, fewer of longer reads
    per_pair = 2 * c["rl"][1]
    e2e_pairs = 0 if args.no_e2e else min(n, max(1_000_000, (4_000_000 * 300 // per_pair) // GEN_CHUNK * GEN_CHUNK))
    for ci, start in enumerate(range(0, n, GEN_CHUNK)):
        cnt = min(GEN_CHUNK, n - start)
        rect = synth.generate_config(args.config, n=cnt, device=dev, chunk_index=rank * 100_000 + ci)
        f_data, f_off, r_data, r_off = rect.to_flat_tensors()
        alg_bytes += int(algorithmic_read_bytes(rect.flen, rect.rlen).sum().item())
        reads, meta, ml, total = ctx.pack_device(f_data, f_off, r_data, r_off)
        meta[:, 0] += base16
        base16 += total // 16
        max_len = max(max_len, ml)
        reads_parts.append(reads[:total])
        meta_parts.append(meta)
        if start < e2e_pairs:
            keep_flat.append(synth.FlatBatch(f_data.cpu().numpy(), f_off.cpu().numpy().astype(np.uint64),
                                             r_data.cpu().numpy(), r_off.cpu().numpy().astype(np.uint64)))
        del rect, f_data, r_data
    reads = torch.cat(reads_parts + [torch.zeros(16, dtype=torch.uint8, device=dev)])
    meta = torch.cat(meta_parts)
    del reads_parts, meta_parts
    seq_stride = (2 * max_len + 15) & ~15
    results = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    seq_nt = torch.empty((n, seq_stride // 2), dtype=torch.uint8, device=dev)
    counters = torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64, device=dev)
    seq_p = torch.empty((n, seq_stride), dtype=torch.float64, device=dev) if args.per_base_p else None
    torch.cuda.synchronize(dev)

    def step():
        ctx.assemble_device(cfg, n, max_len, reads, meta, results, seq_nt, seq_p, seq_stride, counters)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    counters.zero_()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for k in range(args.steps):
        step()
        evs[k + 1].record(stream)
    ctx.synchronize()
    barrier()
    if sampler:
        # the timed region is ~0.1 s, shorter than nvidia-smi's sampling period can resolve: keep the same kernel running
        # for ~1 s more (outside the timing, counters restored afterwards) so the clock record is taken under this load
        saved = counters.clone()
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            step()
            ctx.synchronize()
        counters.copy_(saved)
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["note"] = "sampled every 50 ms over warm-up + timed steps + a 1 s continuation of the same launches"
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n * world * args.steps / (total_ms / 1e3) / 1e6
    # the step is the library's kernels back to back on one stream: its average duration is their summed duration
    kern_ms = float(np.mean(step_ms))
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    # which kernels a step launches and how long each runs: three more steps, outside the timed region, with the library's
    # own events around every kernel (pb_set_timing / pb_last_timing)
    saved = counters.clone()
    ctx.set_timing(True)
    kt = []
    for _ in range(3):
        step()
        kt.append(ctx.last_timing())
    ctx.set_timing(False)
    ctx.synchronize()
    counters.copy_(saved)
    kind = kt[-1][0]
    kms = [float(np.mean([t[1][k] for t in kt])) for k in range(3)]
    fl0 = int(meta[0, 1].item()) & 0xFFFF
    rl0 = (int(meta[0, 1].item()) >> 16) & 0xFFFF
    if kind == 2:
        # algorithmic bytes per pair of each kernel: seeding reads the packed bases + metadata and writes a 32-byte mask record;
        # the lane kernel reads the whole record, the metadata and the mask record and writes the result + the merged read
        seed_b = ((fl0 + 1) // 2 + (rl0 + 1) // 2 + 8) if max_len <= 160 else (alg_bytes / n - (alg_bytes / n - 8) * 2 / 3)
        # batches with reads above 160 nt are listed by length class first (pb::class_list_kernel) and every class that can hold
        # a pair of the batch (<= 160, <= 256, <= 320 nt) runs its own seeding, bin list and lane kernel
        ncls = 1 if (max_len <= 160 or cfg.maxoverlap != 0) else (2 if max_len <= 256 else 3)
        by_class = "" if ncls == 1 else f", once per length class ({ncls} classes; pb::class_list_kernel first)"
        kernels = [
            {"name": ("pbs::sweep_seed_kernel (K1-K3 as a bit-parallel sweep over diagonals, one lane per pair)" if cfg.maxoverlap == 0
                      else "pb::seed_kernel (K1-K3: k-mer join, one warp per pair)") + " + pb::bin_order_kernel (pairs listed by overlap bin)" + by_class,
             "ms": kms[0], "launches_per_step": 2 * ncls + (1 if ncls > 1 else 0),
             "algorithmic_read_bytes_per_pair": seed_b, "achieved_gbs": seed_b * n / (kms[0] / 1e3) / 1e9},
            {"name": "pbl::assemble_lanes_kernel (K4-K6: score + merge, one lane per pair)" + by_class, "ms": kms[1], "launches_per_step": ncls,
             "algorithmic_read_bytes_per_pair": alg_bytes / n + 32, "achieved_gbs": (alg_bytes + 32 * n) / (kms[1] / 1e3) / 1e9},
            {"name": "pb::assemble_kernel, list mode (the pairs the two kernels above hand on)", "ms": kms[2], "launches_per_step": 1},
        ]
        kernel_name = ("pbs::sweep_seed_kernel" if cfg.maxoverlap == 0 else "pb::seed_kernel") + \
            " + pb::bin_order_kernel + pbl::assemble_lanes_kernel + pb::assemble_kernel<list> (one step; the read bytes of the path over their summed duration)"
        launches_per_step = 3 * ncls + 1 + (1 if ncls > 1 else 0)
    else:
        kernels = [{"name": "pb::assemble_kernel", "ms": kms[2], "launches_per_step": 1}]
        kernel_name = "pb::assemble_kernel"
        launches_per_step = 1
    for k in kernels:
        if "achieved_gbs" in k:
            k["frac_of_peak"] = None      # filled in below, once the peak is known
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    for k in kernels:
        if "achieved_gbs" in k:
            k["frac_of_peak"] = k["achieved_gbs"] / peak
    # DRAM traffic of the kernel from the committed `ncu --set full` capture (bytes per pair x pairs of one launch)
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(str(args.config))
        if tr and not args.per_base_p:
            traffic, traffic_src = tr["dram_bytes_per_pair"] * n, tr["source"]
    except Exception:
        pass

    # ---- counters: the STAT merge ---------------------------------------------------------------
    from pandaseq_b200.shard import dist_merge_counters
    ctot = dist_merge_counters(counters, dist) if dist is not None else counters.clone()
    ctot = ctot.cpu().numpy()
    stat = {k: int(ctot[i]) // args.steps for k, i in (("count", pb.C_COUNT), ("ok", pb.C_OK), ("lowq", pb.C_LOWQ),
                                                      ("noalgn", pb.C_NOALGN), ("badr", pb.C_BADR), ("slow", pb.C_SLOW))}

    # ---- e2e: host buffers through pb_assemble_host -----------------------------------------------
    e2e = None
    if not args.no_e2e and keep_flat:
        flat = synth.FlatBatch.concat(keep_flat)
        del keep_flat
        ne = flat.n

        def pinned(arr):
            """host copy of arr in page-locked memory (what a batching caller would read its FASTQ chunks into)"""
            t = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0].copy()).dtype, pin_memory=True)
            t.numpy()[...] = arr
            return t

        keep = [pinned(flat.f_data), pinned(flat.f_off.view(np.int64)), pinned(flat.r_data), pinned(flat.r_off.view(np.int64))]
        flat = synth.FlatBatch(keep[0].numpy(), keep[1].numpy().view(np.uint64), keep[2].numpy(), keep[3].numpy().view(np.uint64))
        res_t = torch.zeros((ne, 32), dtype=torch.uint8, pin_memory=True)
        nt_t = torch.zeros((ne, seq_stride // 2), dtype=torch.uint8, pin_memory=True)
        res_h, nt_h = res_t.numpy(), nt_t.numpy()
        cnt_h = np.zeros(pb.PB_NCOUNTERS, dtype=np.int64)
        code_t = torch.zeros((ne, seq_stride), dtype=torch.int16, pin_memory=True) if args.per_base_p else None
        import ctypes as C
        L = pb.lib()

        def e2e_step():
            if code_t is not None:
                rc = L.pb_assemble_host_codes(ctx._h, C.byref(cfg), ne, flat.f_data.ctypes.data, flat.f_off.ctypes.data, flat.r_data.ctypes.data,
                                              flat.r_off.ctypes.data, res_h.ctypes.data, nt_h.ctypes.data, code_t.data_ptr(), seq_stride, cnt_h.ctypes.data)
            else:
                rc = L.pb_assemble_host(ctx._h, C.byref(cfg), ne, flat.f_data.ctypes.data, flat.f_off.ctypes.data, flat.r_data.ctypes.data,
                                        flat.r_off.ctypes.data, res_h.ctypes.data, nt_h.ctypes.data, None, seq_stride, cnt_h.ctypes.data)
            if rc != 0:
                raise RuntimeError(L.pb_last_error().decode())

        e2e_step()
        barrier()
        ksteps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = flat.f_data.nbytes + flat.r_data.nbytes + flat.f_off.nbytes + flat.r_off.nbytes + 4 * ne
        d2h = res_h.nbytes + nt_h.nbytes + (code_t.numel() * 2 if code_t is not None else 0)
        e2e = {"value": ne * world * ksteps / dt / 1e6, "unit": "Mpairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "pairs_per_step_per_gpu": ne, "steps": ksteps,
               "note": ("pb_assemble_host_codes(): as pb_assemble_host(), plus the per-base log p of every merged base as 16-bit codes into the posterior table (D2H-bound: 792 B per pair out)"
                        if code_t is not None else
                        "pb_assemble_host(): pinned host panda_qual arrays -> H2D -> pack -> assemble -> D2H (results + merged reads) into pinned host arrays, 2 streams")}
        # the same with records the caller keeps in the packed layout (pb_assemble_host_packed): 26 % fewer bytes in, no pack kernel.
        # The packing itself (pb_pack_host, once, outside the timing) is what a parser writing this layout would do instead of AoS.
        reads_p = meta_p = None
        if code_t is None:          # (the packed entry point has no per-base variant)
            reads_h, meta_h, ml_h = pb.pack_host(flat)
            reads_p, meta_p = pinned(reads_h), pinned(meta_h.view(np.int64))
            del reads_h, meta_h

            def e2e_packed_step():
                rc = L.pb_assemble_host_packed(ctx._h, C.byref(cfg), ne, ml_h, reads_p.data_ptr(), meta_p.data_ptr(), res_h.ctypes.data,
                                               nt_h.ctypes.data, seq_stride, cnt_h.ctypes.data)
                if rc != 0:
                    raise RuntimeError(L.pb_last_error().decode())

            e2e_packed_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                e2e_packed_step()
            dtp = time.perf_counter() - t0
            tt = torch.tensor([dtp], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dtp = float(tt.item())
            e2e["packed_input"] = {"value": ne * world * ksteps / dtp / 1e6, "unit": "Mpairs/s",
                                   "h2d_bytes_per_step": int(reads_p.numel() + meta_p.numel() * 8), "d2h_bytes_per_step": int(d2h),
                                   "note": "pb_assemble_host_packed(): pinned host records in the packed layout (4-bit nt + 8-bit PHRED) -> H2D -> assemble -> D2H"}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------
    if e2e is not None:          # the e2e leg's host buffers are done with
        del flat, keep, res_t, nt_t, res_h, nt_h, reads_p, meta_p, code_t
        import gc
        gc.collect()
    cpu = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        probe = synth.generate_config(args.config, n=20_000, device="cpu").to_flat()
        r0, kind = cpu_rate(cfg, probe, cores)
        sample = int(min(max(r0 * 1e6 * 12.0, 20_000), 8_000_000 * 300 // per_pair))
        flat_c = synth.generate_config(args.config, n=sample, device="cpu", chunk_index=1).to_flat()
        r1, kind = cpu_rate(cfg, flat_c, cores)
        cpu = {"value": r1, "unit": "Mpairs/s", "cores": cores, "kind": kind,
               "sample": f"{sample} synthetic pairs of the same config, panda_assembler_assemble loop, one assembler per thread, logging off"}
    if dist is not None:
        dist.barrier()          # the other ranks wait for rank 0's CPU leg, so that all leave together

    if rank == 0:
        line = {
            "metric": "read-pairs/s (Mpairs/s), pair assembly hot path", "value": value, "unit": "Mpairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 (4-bit nt, 8-bit PHRED) + f64 log-probabilities", "data": "synthetic",
            "config": {"workload": f"BASELINE config {args.config}: {n} synthetic {shape_label(c)} pairs per GPU, {c['algo']}"
                                   + (", primer strip" if kw else ""),
                       "pairs_per_gpu": n, "l2_policy": f"inputs larger than L2 ({reads.numel() / 1e6:.0f} MB packed per GPU), no flush",
                       "outputs": "32 B result record + merged read (4 bit/base) per pair; "
                                  + ("per-base log p of every merged base as well (f64 on the device-resident leg, 16-bit codes into the posterior table on the e2e leg)"
                                     if args.per_base_p else "per-base log p not requested"),
                       "parallelism": f"{world} x independent shards, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_pair": alg_bytes / n,
                         "kernel": kernel_name, "kernel_ms": kern_ms, "kernels": kernels},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks, "stat": stat,
            "host_affinity": affinity,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
