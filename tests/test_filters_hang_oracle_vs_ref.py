"""The overhang trimmer (hang.c) and the post-assembly filters (module.c checks) of the oracle against the compiled
reference: the reference's own panda_trim_overhangs / no_n_check / short_check / long_check and its min_phred,
min_overlapbits and completely_miss_the_point plugins run inside the harness."""
import numpy as np
import pytest

import datasets
import oracle_lib
import pandaseq_b200 as pb
from test_oracle_vs_ref import same

pytestmark = pytest.mark.skipif(not oracle_lib.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("filters", datasets.FILTER_SETS)
@pytest.mark.parametrize("algo", ["simple_bayesian", "rdp_mle"])
def test_filters(built, filters, algo):
    b = datasets.cfg1(1200)
    cfg = pb.make_config(algo, filters=filters)
    a, r = oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b)
    same(a, r)
    if algo == "rdp_mle" or not any(name == "min_overlapbits" for name, _ in filters):      # (that plugin only makes sense for log-odds scores)
        assert 0 < (a["status"] >= 8).sum() < len(a["status"]), "the filter set should reject some pairs, not all"
    assert a["counters"][pb.C_REJECTED:pb.C_REJECTED + 7].sum() == (a["status"] >= 8).sum()


def test_filters_after_primer_strip(built):
    fwd, rev = datasets.primer_codes()
    b = datasets.primers300(600)
    cfg = pb.make_config("simple_bayesian", forward_primer=fwd, reverse_primer=rev, post_primers=True,
                         filters=[("short", 440), ("min_phred", 3), ("long", 480)])
    same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))


@pytest.mark.parametrize("skip", [False, True])
@pytest.mark.parametrize("which", ["both", "forward", "reverse"])
def test_overhang_trimmer(built, skip, which):
    hf, hr = datasets.overhang_codes()
    b = datasets.overhang(600)
    kw = dict(hang_forward=hf if which != "reverse" else None, hang_reverse=hr if which != "forward" else None, hang_skip=skip)
    for algo in ("simple_bayesian", "pear"):
        cfg = pb.make_config(algo, **kw)
        a, r = oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b)
        same(a, r)
        plain = oracle_lib.assemble("port", pb.make_config(algo), b)
        assert (a["seq_len"] != plain["seq_len"]).sum() > 50, "trimming should change many assemblies"


def test_overhang_with_filters_and_thresholds(built):
    hf, hr = datasets.overhang_codes()
    b = datasets.overhang(600, seed=4)
    cfg = pb.make_config("simple_bayesian", hang_forward=hf, hang_reverse=hr, hang_threshold=np.log(0.8), filters=[("short", 80), ("no_n", 0)])
    same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))
    # offset.c:72 divides by index + 1, so at ordinary thresholds the scan "finds" the sequence in every read; a threshold this
    # strict makes the trimmer drop pairs (status 6 = never reached the assembler, not counted)
    cfg = pb.make_config("simple_bayesian", hang_forward=hf, hang_reverse=hr, hang_threshold=np.log(0.9997))
    a, r = oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b)
    same(a, r)
    assert 0 < (a["status"] == 6).sum() < len(a["status"])
    assert a["counters"][pb.C_COUNT] == (a["status"] != 6).sum()
