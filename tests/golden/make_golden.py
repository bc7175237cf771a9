#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_harness.so, compiled from
/root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds a small seeded input batch (flat panda_qual arrays), the configuration, and what
panda_assembler_assemble() returned for every pair (status, overlap, merged bases, per-base log p, quality,
counters ...).  The inputs are stored too, so the fixtures do not depend on the random number generator
of whatever torch version later reads them.  tests/test_golden.py checks the oracle port (CPU) and
tests/test_gpu_golden.py the CUDA path against these files."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
import oracle_lib  # noqa: E402
import pandaseq_b200 as pb  # noqa: E402

CASES = {
    "cfg1_simple_bayesian": (lambda: datasets.cfg1(600), dict(algo="simple_bayesian")),
    "cfg1_pear": (lambda: datasets.cfg1(600), dict(algo="pear")),
    "cfg1_rdp_mle": (lambda: datasets.cfg1(600), dict(algo="rdp_mle")),
    "cfg1_flash": (lambda: datasets.cfg1(600), dict(algo="flash")),
    "cfg1_ea_util": (lambda: datasets.cfg1(400), dict(algo="ea_util")),
    "cfg1_stitch": (lambda: datasets.cfg1(400), dict(algo="stitch")),
    "cfg1_uparse": (lambda: datasets.cfg1(400), dict(algo="uparse")),
    "stress_stitch_maxov300": (lambda: datasets.stress(300), dict(algo="stitch", maxoverlap=300)),
    "edge_cases_ea_util_maxov800": (lambda: datasets.edge_cases(), dict(algo="ea_util", maxoverlap=800)),
    "stress_sb_maxov300": (lambda: datasets.stress(400), dict(algo="simple_bayesian", maxoverlap=300)),
    "stress_pear_trims": (lambda: datasets.stress(400, seed=3), dict(algo="pear", forward_trim=20, reverse_trim=20, maxoverlap=300)),
    "mixed_sb": (lambda: datasets.mixed(400), dict(algo="simple_bayesian")),
    "long250_pear": (lambda: datasets.long250(200), dict(algo="pear")),
    "primers300_rdp": (lambda: datasets.primers300(150), dict(algo="rdp_mle", primers=True)),
    "primers300_rdp_penalty": (lambda: datasets.primers300(150), dict(algo="rdp_mle", primers=True, primer_penalty=0.0005)),
    "primers300_after_sb": (lambda: datasets.primers300(200), dict(algo="simple_bayesian", primers=True, post_primers=True)),
    "cfg1_after_trims_pear": (lambda: datasets.cfg1(300), dict(algo="pear", post_primers=True, forward_trim=12, reverse_trim=7)),
    "lowcomplexity_sb": (lambda: datasets.low_complexity(200), dict(algo="simple_bayesian")),
    "edge_cases_sb": (lambda: datasets.edge_cases(), dict(algo="simple_bayesian")),
    "edge_cases_rdp_maxov800": (lambda: datasets.edge_cases(), dict(algo="rdp_mle", maxoverlap=800)),
    "cfg1_filters_rdp": (lambda: datasets.cfg1(300), dict(algo="rdp_mle", filters=datasets.FILTER_SETS[-1])),
    "cfg1_filters_sb": (lambda: datasets.cfg1(300), dict(algo="simple_bayesian", filters=[("long", 230), ("no_n", 0), ("short", 190), ("min_phred", 8)])),
    # plugin_pear_test.c at the only parameters the reference can run it with (its defaults, see oracle/ref_harness.c)
    "stress_pear_test_sb": (lambda: datasets.stress(400), dict(algo="simple_bayesian", filters=[("pear_test", (1.0, -1.0, 0.01))])),
    "mixed_pear_test_filters_sb": (lambda: datasets.mixed(300), dict(algo="simple_bayesian", filters=[("short", 120), ("pear_test", (1.0, -1.0, 0.01)), ("long", 500)])),
    "overhang_sb": (lambda: datasets.overhang(300), dict(algo="simple_bayesian", hang=True)),
    "overhang_strict_pear_filters": (lambda: datasets.overhang(300, seed=4), dict(algo="pear", hang=True, hang_threshold=-0.0003, filters=[("short", 80)])),
}


def build_config(spec):
    kw = dict(spec)
    algo = kw.pop("algo")
    if kw.pop("primers", False):
        fwd, rev = datasets.primer_codes()
        kw.update(forward_primer=fwd, reverse_primer=rev)
    if kw.pop("hang", False):
        hf, hr = datasets.overhang_codes()
        kw.update(hang_forward=hf, hang_reverse=hr)
    return pb.make_config(algo, **kw)


def main():
    if not oracle_lib.have_ref():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` (needs /root/reference)")
    only = set(sys.argv[1:])          # `make_golden.py name ...` regenerates just those fixtures
    for name, (mk, spec) in CASES.items():
        if only and name not in only:
            continue
        batch = mk()
        cfg = build_config(spec)
        out = oracle_lib.assemble("ref", cfg, batch)
        ok = out["status"] == 0
        width = int(out["seq_len"].max()) if ok.any() else 0
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            f_data=batch.f_data, f_off=batch.f_off, r_data=batch.r_data, r_off=batch.r_off,
            spec=np.array(repr(spec)),
            status=out["status"], slow=out["slow"], overlap=out["overlap"], seq_len=out["seq_len"],
            mismatches=out["mismatches"], degenerates=out["degenerates"], examined=out["examined"],
            fwd_offset=out["fwd_offset"], rev_offset=out["rev_offset"], quality=out["quality"], est_prob=out["est_prob"],
            seq_nt=out["seq_nt"][:, :width], seq_p=out["seq_p"][:, :width], counters=out["counters"])
        print(f"{name}: {batch.n} pairs, {int(ok.sum())} assembled, width {width}")


if __name__ == "__main__":
    main()
