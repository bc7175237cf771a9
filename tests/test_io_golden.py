"""The oracle's FASTQ reader, identifier parser, PHRED search and FASTA/FASTQ writer against the golden vectors generated
from the compiled reference (tests/golden/make_golden_io.py).  CPU only; runs where the reference build is absent."""
import ast
import os

import numpy as np
import pytest

import oracle_lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FQ = np.load(os.path.join(GOLD, "io_fastq.npz"))
FM = np.load(os.path.join(GOLD, "io_format.npz"))
ID_FIELDS = ("instrument", "run", "flowcell", "lane", "tile", "x", "y", "tag")


def golden_case(name):
    kw = ast.literal_eval(str(FQ[f"{name}.kw"]))
    return bytes(FQ[f"{name}.fwd"]), bytes(FQ[f"{name}.rev"]), kw


def check_against_golden(name, got):
    assert got["n"] == int(FQ[f"{name}.n"])
    assert got["error"] == int(FQ[f"{name}.error"])
    for k in ID_FIELDS:
        assert np.array_equal(got["ids"][k], FQ[f"{name}.ids"][k]), k
    for k in ("f_data", "f_off", "r_data", "r_off"):
        assert np.array_equal(getattr(got["batch"], k), FQ[f"{name}.{k}"]), k


@pytest.mark.parametrize("name", [str(n) for n in FQ["names"]])
def test_fastq_golden(name):
    f, r, kw = golden_case(name)
    check_against_golden(name, oracle_lib.fastq_parse("port", f, r, **kw))


def test_seqid_golden():
    for policy, text, rc, fmt, want in zip(FQ["hdr.policy"], FQ["hdr.text"], FQ["hdr.rc"], FQ["hdr.fmt"], FQ["hdr.ids"]):
        got_rc, got_fmt, got = oracle_lib.seqid_parse("port", bytes(text), int(policy))
        assert got_rc == rc, text
        if rc:
            assert got_fmt == fmt
            for k in ID_FIELDS:
                assert got[k] == want[k], (text, k)


def test_phred_golden():
    for p, want in zip(FM["phred_p"], FM["phred"]):
        assert oracle_lib.result_phred("port", float(p)) == want


@pytest.mark.parametrize("fastq", [False, True])
def test_format_golden(fastq):
    width = FM["seq_nt"].shape[1]
    text = oracle_lib.format_flat("port", fastq, FM["ids"], FM["status"], FM["quality"], FM["seq_len"], FM["seq_nt"], FM["seq_p"], width)
    assert text == bytes(FM["fastq" if fastq else "fasta"])
