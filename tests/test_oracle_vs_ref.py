"""Pins the oracle: oracle/panda_oracle.c against the compiled reference, every field bit-identical.
Runs where oracle/_ref exists (built in the container that has /root/reference; it travels to the GPU box)."""
import numpy as np
import pytest

import datasets
import oracle_lib
import pandaseq_b200 as pb

pytestmark = pytest.mark.skipif(not oracle_lib.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def same(a, r):
    ok = r["status"] == 0
    assert np.array_equal(a["status"], r["status"]) and np.array_equal(a["slow"], r["slow"])
    assert np.array_equal(a["counters"], r["counters"])
    for k in ("overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset"):
        assert np.array_equal(a[k][ok], r[k][ok]), k
    for k in ("quality", "est_prob"):
        assert np.array_equal(a[k][ok].view(np.uint64), r[k][ok].view(np.uint64)), k
    assert np.array_equal(a["seq_nt"][ok], r["seq_nt"][ok])
    assert np.array_equal(a["seq_p"][ok].view(np.uint64), r["seq_p"][ok].view(np.uint64))


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle", "flash", "ea_util", "stitch", "uparse"])
def test_cfg1(built, algo):
    b = datasets.cfg1(4000)
    cfg = pb.make_config(algo)
    same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))


@pytest.mark.parametrize("kw", [dict(), dict(maxoverlap=300), dict(forward_trim=20, reverse_trim=20, maxoverlap=300),
                                dict(threshold=0.9), dict(minoverlap=30, threshold=0.3), dict(sb_q=0.1)])
def test_stress_options(built, kw):
    b = datasets.stress(2000)
    for algo in ("simple_bayesian", "pear", "rdp_mle"):
        cfg = pb.make_config(algo, **kw)
        same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))


def test_mixed_lengths_and_long_reads(built):
    for b in (datasets.mixed(2000), datasets.long250(800)):
        for algo in ("simple_bayesian", "pear", "rdp_mle"):
            cfg = pb.make_config(algo)
            same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))


@pytest.mark.parametrize("penalty", [0.0, 0.0005])
def test_primers(built, penalty):
    fwd, rev = datasets.primer_codes()
    b = datasets.primers300(800)
    cfg = pb.make_config("rdp_mle", forward_primer=fwd, reverse_primer=rev, primer_penalty=penalty)
    same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))
    # the scan on its own, forward and reverse direction
    rng = np.random.default_rng(0)
    for i in range(0, 200, 7):
        f, r = b.pair(i)
        for read in (f, r):
            for rv in (False, True):
                needle = bytes(fwd.tolist())
                assert oracle_lib.compute_offset("port", cfg.threshold, penalty, rv, read, needle) == \
                    oracle_lib.compute_offset("ref", cfg.threshold, penalty, rv, read, needle)


def test_low_complexity_and_edge_cases(built):
    for b in (datasets.low_complexity(), datasets.edge_cases()):
        for algo in ("simple_bayesian", "pear", "rdp_mle", "flash", "ea_util", "stitch", "uparse"):
            for kw in (dict(), dict(maxoverlap=800)):
                cfg = pb.make_config(algo, **kw)
                same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))


def test_threads_do_not_change_results(built):
    b = datasets.cfg1(3000)
    cfg = pb.make_config("simple_bayesian")
    same(oracle_lib.assemble("port", cfg, b, threads=4), oracle_lib.assemble("ref", cfg, b, threads=3))


def test_primers_after(built):
    fwd, rev = datasets.primer_codes()
    for b, kw in ((datasets.primers300(600), dict(forward_primer=fwd, reverse_primer=rev)),
                  (datasets.primers300(300), dict(forward_primer=fwd, reverse_primer=rev, primer_penalty=0.0005)),
                  (datasets.cfg1(800), dict(forward_trim=12, reverse_trim=7)),
                  (datasets.stress(800), dict(forward_primer=fwd[:6], reverse_primer=rev[:5], maxoverlap=300))):
        for algo in ("simple_bayesian", "rdp_mle"):
            cfg = pb.make_config(algo, post_primers=True, **kw)
            same(oracle_lib.assemble("port", cfg, b), oracle_lib.assemble("ref", cfg, b))


def test_many_threads_many_calls(built):
    """The reference arm of bench.py calls the harness once per step with one assembler per host thread.  Every assembler
    owns a PandaWriter, every PandaWriter a pthread_key (writer.c:93), and a process has 1024 keys: a harness that leaks
    writers dies after 1024 / (2 x threads) calls (round 1: SIGSEGV on the 32-core box).  64 threads x 40 calls."""
    b = datasets.cfg1(640)
    cfg = pb.make_config("simple_bayesian")
    first = oracle_lib.assemble("ref", cfg, b, threads=64)
    for _ in range(39):
        again = oracle_lib.assemble("ref", cfg, b, threads=64)
    same(first, again)
    same(oracle_lib.assemble("port", cfg, b), again)
