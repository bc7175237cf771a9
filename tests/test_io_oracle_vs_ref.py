"""Pins oracle/panda_oracle_io.c (FASTQ parse, identifier parse, output formatting, PHRED search) against the
compiled reference (oracle/_ref).  Skipped where the reference build is absent; the same cases are then covered by
tests/golden/io_*.npz (test_io_golden.py)."""
import numpy as np
import pytest

import oracle_lib
from fastq_cases import HEADERS, REF_UNDEFINED, file_cases, rng_garbage
from pandaseq_b200 import make_config, synth

pytestmark = pytest.mark.skipif(not oracle_lib.have_ref(), reason="reference build (oracle/_ref) not present")


def _same_parse(a, b, check_error=True):
    assert a["n"] == b["n"]
    if check_error:
        assert a["error"] == b["error"], (oracle_lib.FQ_ERRORS[a["error"]], oracle_lib.FQ_ERRORS[b["error"]])
    for k in ("instrument", "run", "flowcell", "lane", "tile", "x", "y", "tag"):
        assert np.array_equal(a["ids"][k], b["ids"][k]), k
    for k in ("f_data", "f_off", "r_data", "r_off"):
        assert np.array_equal(getattr(a["batch"], k), getattr(b["batch"], k)), k


@pytest.mark.parametrize("policy", [0, 1, 2])
def test_seqid_parse(policy):
    for h in HEADERS:
        rp, fp, ip = oracle_lib.seqid_parse("port", h, policy)
        rr, fr, ir = oracle_lib.seqid_parse("ref", h, policy)
        if len(h.split(b":")[0]) > 100 or (rr == 0 and rp == 0):
            assert rp == 0 or rr == rp      # over-long fields: refused here, UB there (documented)
            continue
        assert (rp, fp) == (rr, fr), h
        for k in ("instrument", "run", "flowcell", "lane", "tile", "x", "y", "tag"):
            assert ip[k] == ir[k], (h, k)


@pytest.mark.parametrize("name", sorted(file_cases()))
def test_fastq_cases(name):
    f, r, kw = file_cases()[name]
    if name in REF_UNDEFINED:
        pytest.skip("undefined behaviour in the reference (reads past its line buffer)")
    _same_parse(oracle_lib.fastq_parse("port", f, r, **kw), oracle_lib.fastq_parse("ref", f, r, **kw))


def test_fastq_small_reads_from_source():
    """linebuf refills: the source hands out 7 bytes at a time"""
    f, r, kw = file_cases()["clean"]
    _same_parse(oracle_lib.fastq_parse("port", f, r), oracle_lib.fastq_parse("ref", f, r, max_read=7))


@pytest.mark.parametrize("seed", range(6))
def test_fastq_garbage(seed):
    f, r = rng_garbage(seed), rng_garbage(seed + 100)
    _same_parse(oracle_lib.fastq_parse("port", f, r, policy=2), oracle_lib.fastq_parse("ref", f, r, policy=2))
    # a good file with one garbage record spliced in
    gf, gr, _ = file_cases()["clean"]
    cut = gf.index(b"\n@", len(gf) // 2) + 1
    _same_parse(oracle_lib.fastq_parse("port", gf[:cut] + f, gr), oracle_lib.fastq_parse("ref", gf[:cut] + f, gr))


def test_result_phred():
    t = oracle_lib.tables("port")
    ps = list(t["score"]) + list(np.nextafter(t["score"], 0)) + list(np.nextafter(t["score"], -10)) + [-5.0, -2.0, -1e-9, 0.0, 0.5]
    ps += list(np.random.default_rng(3).uniform(-3, 0, 500))
    for tabname in ("match_sb", "mismatch_rdp_asm", "match_pear"):
        ps += list(t[tabname].reshape(-1)[::37])
    for p in ps:
        assert oracle_lib.result_phred("port", p) == oracle_lib.result_phred("ref", p), p


@pytest.mark.parametrize("fastq", [False, True])
@pytest.mark.parametrize("algo", ["simple_bayesian", "rdp_mle"])
def test_format(fastq, algo):
    b = synth.generate_config(1, n=300, n_rate=0.01, btail_rate=0.1).to_flat()
    f, r = synth.fastq_pair(b)
    parsed = oracle_lib.fastq_parse("ref", bytes(f.numpy()), bytes(r.numpy()))
    res = oracle_lib.assemble("ref", make_config(algo), parsed["batch"])
    args = (fastq, parsed["ids"], res["status"], res["quality"], res["seq_len"], res["seq_nt"], res["seq_p"], res["seq_stride"])
    a, bb = oracle_lib.format_flat("port", *args), oracle_lib.format_flat("ref", *args)
    assert len(a) > 1000 and a == bb
