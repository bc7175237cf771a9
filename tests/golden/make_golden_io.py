#!/usr/bin/env python
"""Generate tests/golden/io_fastq.npz and tests/golden/io_format.npz from the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden_io.py

io_fastq.npz   every FASTQ case of tests/fastq_cases.py that the reference defines (the texts themselves, the arguments, and
               what panda_create_fastq_reader's PandaNextSeq delivered: pairs, identifiers, the code it logged), plus the
               identifier corpus through panda_seqid_parse_fail.
io_format.npz  300 assembled pairs (panda_assembler_assemble, simple_bayesian) and the text panda_output_fasta /
               panda_output_fastq wrote for them, plus panda_result_phred over a grid of probabilities.
tests/test_io_golden.py checks the oracle port (CPU) and tests/test_gpu_io.py the CUDA path against these files."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
import pandaseq_b200 as pb  # noqa: E402
from fastq_cases import HEADERS, REF_UNDEFINED, file_cases  # noqa: E402
from pandaseq_b200 import synth  # noqa: E402


def main():
    if not oracle_lib.have_ref():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` (needs /root/reference)")
    out = {}
    names = []
    for name, (f, r, kw) in sorted(file_cases().items()):
        if name in REF_UNDEFINED:
            continue
        res = oracle_lib.fastq_parse("ref", f, r, **kw)
        names.append(name)
        out[f"{name}.fwd"] = np.frombuffer(f, dtype=np.uint8)
        out[f"{name}.rev"] = np.frombuffer(r, dtype=np.uint8)
        out[f"{name}.kw"] = np.array(repr(kw))
        out[f"{name}.n"] = np.int64(res["n"])
        out[f"{name}.error"] = np.int64(res["error"])
        out[f"{name}.ids"] = res["ids"]
        for k in ("f_data", "f_off", "r_data", "r_off"):
            out[f"{name}.{k}"] = getattr(res["batch"], k)
    out["names"] = np.array(names)
    hdr = []
    for policy in (0, 1, 2):
        for h in HEADERS:
            if len(h.split(b":")[0]) > 100:
                continue        # a 101+ character field overruns the reference's struct member (seqid.c:153): undefined
            rc, fmt, ident = oracle_lib.seqid_parse("ref", h, policy)
            hdr.append((policy, h, rc, fmt, ident))
    out["hdr.policy"] = np.array([x[0] for x in hdr])
    out["hdr.text"] = np.array([x[1] for x in hdr], dtype="S200")
    out["hdr.rc"] = np.array([x[2] for x in hdr])
    out["hdr.fmt"] = np.array([x[3] for x in hdr])
    out["hdr.ids"] = np.array([x[4] for x in hdr], dtype=oracle_lib.SEQID_DTYPE)
    np.savez_compressed(os.path.join(HERE, "io_fastq.npz"), **out)
    print(f"io_fastq.npz: {len(names)} files, {len(hdr)} identifiers")

    b = synth.generate_config(1, n=300, n_rate=0.01, btail_rate=0.1, chunk_index=5).to_flat()
    f, r = (bytes(t.numpy()) for t in synth.fastq_pair(b))
    parsed = oracle_lib.fastq_parse("ref", f, r)
    res = oracle_lib.assemble("ref", pb.make_config("simple_bayesian"), parsed["batch"])
    width = int(res["seq_len"].max())
    args = (parsed["ids"], res["status"], res["quality"], res["seq_len"], res["seq_nt"], res["seq_p"], res["seq_stride"])
    t = oracle_lib.tables("ref")
    ps = np.concatenate([t["score"], np.nextafter(t["score"], 0), np.nextafter(t["score"], -10), np.linspace(-3, 0, 400),
                         t["match_sb"].reshape(-1)[::29], t["mismatch_rdp_asm"].reshape(-1)[::31]])
    np.savez_compressed(os.path.join(HERE, "io_format.npz"), fwd=np.frombuffer(f, dtype=np.uint8), rev=np.frombuffer(r, dtype=np.uint8),
                        ids=parsed["ids"], status=res["status"], quality=res["quality"], seq_len=res["seq_len"],
                        seq_nt=res["seq_nt"][:, :width], seq_p=res["seq_p"][:, :width],
                        fasta=np.frombuffer(oracle_lib.format_flat("ref", False, *args), dtype=np.uint8),
                        fastq=np.frombuffer(oracle_lib.format_flat("ref", True, *args), dtype=np.uint8),
                        phred_p=ps, phred=np.array([oracle_lib.result_phred("ref", p) for p in ps], dtype=np.int8))
    print(f"io_format.npz: {int((res['status'] == 0).sum())} assembled pairs")


if __name__ == "__main__":
    main()
