"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Integer/base outputs bit-exact, floating point within 1e-6 (BASELINE.json north_star)."""
import numpy as np
import pytest

import datasets
import oracle_lib
import pandaseq_b200 as pb
from parity import compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = pb.Context(0)
    yield c
    c.close()


def run_both(ctx, cfg, batch):
    got = ctx.assemble_host(cfg, batch, want_nt=True, want_p=True)
    want = oracle_lib.assemble("port", cfg, batch)
    return got, want, compare(got, want)


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle", "flash"])
def test_cfg1_all_algorithms(ctx, algo):
    got, want, rep = run_both(ctx, pb.make_config(algo), datasets.cfg1())
    assert rep["ok"], rep
    assert rep["max_dq"] <= 1e-9 and rep["max_dp"] <= 1e-9


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle"])
@pytest.mark.parametrize("maxoverlap", [0, 300])
def test_read_through_stress(ctx, algo, maxoverlap):
    got, want, rep = run_both(ctx, pb.make_config(algo, maxoverlap=maxoverlap), datasets.stress())
    assert rep["ok"], rep


def test_trims_and_thresholds(ctx):
    b = datasets.stress(2000, seed=3)
    for kw in (dict(forward_trim=20, reverse_trim=20, maxoverlap=300), dict(forward_trim=7, reverse_trim=0),
               dict(threshold=0.9), dict(threshold=0.3, minoverlap=30), dict(minoverlap=140, maxoverlap=145), dict(sb_q=0.1)):
        got, want, rep = run_both(ctx, pb.make_config("simple_bayesian", **kw), b)
        assert rep["ok"], (kw, rep)


@pytest.mark.parametrize("algo", ["simple_bayesian", "rdp_mle"])
def test_mixed_lengths(ctx, algo):
    got, want, rep = run_both(ctx, pb.make_config(algo), datasets.mixed())
    assert rep["ok"], rep


def test_mixed_lengths_pear_forward_longer(ctx):
    # pear indexes forward qualities with the reverse index (algo_pear.c:52); defined for R > F as quality 0
    got, want, rep = run_both(ctx, pb.make_config("pear"), datasets.mixed(2000))
    assert rep["ok"], rep


def test_2x250_pear(ctx):
    got, want, rep = run_both(ctx, pb.make_config("pear"), datasets.long250())
    assert rep["ok"], rep


@pytest.mark.parametrize("penalty", [0.0, 0.0005])
def test_primers_rdp_mle(ctx, penalty):
    fwd, rev = datasets.primer_codes()
    cfg = pb.make_config("rdp_mle", forward_primer=fwd, reverse_primer=rev, primer_penalty=penalty)
    got, want, rep = run_both(ctx, cfg, datasets.primers300())
    assert rep["ok"], rep


def test_primers_absent_and_partial(ctx):
    fwd, rev = datasets.primer_codes()
    cfg = pb.make_config("simple_bayesian", forward_primer=fwd, reverse_primer=rev)
    got, want, rep = run_both(ctx, cfg, datasets.cfg1(2000))      # reads without the primers
    assert rep["ok"], rep
    cfg = pb.make_config("simple_bayesian", forward_primer=fwd)    # only a forward primer, reverse trim
    got, want, rep = run_both(ctx, cfg, datasets.primers300(500))
    assert rep["ok"], rep


def test_low_complexity(ctx):
    for algo in ("simple_bayesian", "pear"):
        got, want, rep = run_both(ctx, pb.make_config(algo), datasets.low_complexity())
        assert rep["ok"], rep


def test_edge_cases(ctx):
    b = datasets.edge_cases()
    for algo in ("simple_bayesian", "pear", "rdp_mle", "flash"):
        for kw in (dict(), dict(maxoverlap=800), dict(minoverlap=10)):
            got, want, rep = run_both(ctx, pb.make_config(algo, **kw), b)
            assert rep["ok"], (algo, kw, rep)


def test_empty_batch(ctx):
    b = datasets.cfg1(10).slice(0, 0)
    got = ctx.assemble_host(pb.make_config("simple_bayesian"), b)
    assert len(got["results"]) == 0 and got["counters"].sum() == 0


def test_unsupported_configurations_fail_loudly(ctx):
    b = datasets.cfg1(10)
    with pytest.raises(pb.PandaseqError):
        ctx.assemble_host(pb.make_config("simple_bayesian", num_kmers=3), b)
    with pytest.raises(pb.PandaseqError):
        ctx.assemble_host(pb.make_config("simple_bayesian", post_primers=True), b)
    with pytest.raises(pb.PandaseqError):
        ctx.assemble_host(pb.make_config(7), b)
