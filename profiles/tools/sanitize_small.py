#!/usr/bin/env python
"""A few thousand pairs through every kernel of the pair path (two-kernel path, general kernel, primer scans, overhang trimmer,
primers-after), small enough to run under compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets
import pandaseq_b200 as pb
from pandaseq_b200 import synth

ctx = pb.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
fwd, rev = datasets.primer_codes()
hf, hr = datasets.overhang_codes()
runs = [
    ("two-kernel path", pb.make_config("simple_bayesian"), datasets.cfg1(n), False),
    ("two-kernel path, flash", pb.make_config("flash"), datasets.stress(n), False),
    ("general, per-base p", pb.make_config("simple_bayesian"), datasets.cfg1(n), True),
    ("pear 2x250", pb.make_config("pear"), datasets.long250(n // 2), True),
    ("rdp_mle + primers", pb.make_config("rdp_mle", forward_primer=fwd, reverse_primer=rev), datasets.primers300(n // 2), False),
    ("primers after", pb.make_config("simple_bayesian", forward_primer=fwd, reverse_primer=rev, post_primers=True), datasets.primers300(n // 2), True),
    ("overhang trimmer", pb.make_config("simple_bayesian", hang_forward=hf, hang_reverse=hr, hang_skip=True), datasets.overhang(n // 2), False),
    ("mixed lengths", pb.make_config("simple_bayesian"), datasets.mixed(n // 2), False),
    ("edge cases", pb.make_config("simple_bayesian"), datasets.edge_cases(), False),
    # round 2: the sweep + lane kernels of every length class (class lists, sweep<5/8/10>, lanes<160/256/320>), pear on the lane kernel
    ("length classes 75-300", pb.make_config("simple_bayesian"), synth.generate(n, rl=(75, 300), tmpl=None, seed=91, mixed=True, n_rate=0.0005, btail_rate=0.05).to_flat(), False),
    ("length classes, pear", pb.make_config("pear"), synth.generate(n // 2, rl=(75, 300), tmpl=None, seed=92, mixed=True).to_flat(), False),
    ("pear on the lane kernel", pb.make_config("pear"), datasets.cfg1(n), False),
    ("2x300", pb.make_config("uparse"), synth.generate(n // 2, rl=(300, 300), tmpl=(320, 580), seed=93).to_flat(), False),
]
for name, cfg, batch, want_p in runs:
    got = ctx.assemble_host(cfg, batch, want_nt=True, want_p=want_p)
    print(name, "ok" if got["counters"][pb.C_COUNT] > 0 else "no pairs", int(got["counters"][pb.C_OK]), "assembled", flush=True)
print("lanes stats", ctx.lanes_stats())
ctx.close()
