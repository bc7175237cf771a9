"""FASTQ / identifier edge cases shared by the oracle-vs-reference test, the golden fixture generator and the GPU tests."""
from __future__ import annotations

import numpy as np

from pandaseq_b200 import synth

HEADERS = [
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 1:N:0:1",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 2:N:0:1",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 1:N:0:ACGTAC",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 1:N:0:",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 1:N:0",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 0:N:0:1",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338 1:Y:18:ATCACG",
    b"M01271:10:000000000-A3WGH:001:01101:0015589:1338 1:N:0:1",
    b"M01271:10:000000000-A3WGH:1:1101:15589:13x8 1:N:0:1",
    b"M01271::000000000-A3WGH:1:1101:15589:1338 1:N:0:1",
    b":10:FC:1:2:3:4 1:N:0:1",
    b"HWUSI-EAS100R:6:73:941:1973#0/1",
    b"HWUSI-EAS100R:6:73:941:1973#ACGT/2",
    b"HWUSI-EAS100R:6:73:941:1973/1",
    b"HWUSI-EAS100R:6:73:941:1973#/1",
    b"HWUSI-EAS100R:6:73:941:1973#0/",
    b"HWUSI-EAS100R:6:73:941:1973#0",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338#ACGT/1",
    b"M01271:10:000000000-A3WGH:1:1101:15589:1338/2",
    b"SRR001666.1 071112_SLXA-EAS1_s_7:5:1:817:345 length=36",
    b"ERR12345.678 something",
    b"SRR001666.1",
    b"SRR001666.1 ",
    b"SRR001666.1 x",
    b"SRR00x666.1 abc",
    b"SRR",
    b"SRRx",
    b"",
    b"x",
    b"a:b:c:1:2:3:4 1:N:0:" + b"T" * 50,
    b"a:b:c:1:2:3:4 1:N:0:" + b"T" * 51,
    b"I" * 100 + b":b:c:1:2:3:4 1:N:0:1",
    b"I" * 102 + b":b:c:1:2:3:4 1:N:0:1",
    b"a:b:c:99999999999:2:3:4 1:N:0:1",
    b"a b#c/1:2:3:4 1:N:0:1",
    b"a:b:c:1:2:3:4:1:N:0:1",
    b"a:b:c:1:2:3:4 1:N:0:1 extra",
    b"a:b:c:1:2:3:4 1:N:0:1:more",
]


def records(headers_f, headers_r, seqs_f, quals_f, seqs_r, quals_r, eol=b"\n"):
    f = b"".join(b"@" + h + eol + s + eol + b"+" + eol + q + eol for h, s, q in zip(headers_f, seqs_f, quals_f))
    r = b"".join(b"@" + h + eol + s + eol + b"+" + eol + q + eol for h, s, q in zip(headers_r, seqs_r, quals_r))
    return f, r


def _good(n, seed=11, **kw):
    b = synth.generate_config(1, n=n, n_rate=0.01, btail_rate=0.1, chunk_index=seed).to_flat()
    f, r = synth.fastq_pair(b, **kw)
    return bytes(f.numpy()), bytes(r.numpy())


def _replace_line(text: bytes, record: int, line: int, new: bytes) -> bytes:
    lines = text.split(b"\n")
    lines[4 * record + line] = new
    return b"\n".join(lines)


# linebuf.c:72-77 reads past its data when the last line lacks its '\n' (stale bytes decide; the reference can crash):
# these cases are only checked between the oracle and the device path, whose rule is "an unterminated line is not delivered".
REF_UNDEFINED = {"no_trailing_newline", "truncated_mid_line", "reverse_truncated_mid_line"}


def file_cases():
    """name -> (forward text, reverse text, kwargs for fastq_parse)"""
    cases = {}
    f, r = _good(40)
    cases["clean"] = (f, r, {})
    cases["crlf"] = _good(12, crlf=True) + ({},)
    cases["phred64"] = _good(12, qual_offset=64) + (dict(qualmin=64),)
    cases["phred64_read_as_33"] = _good(12, qual_offset=64) + (dict(qualmin=33),)     # fastq.c:44 clamp quirk: > 33+46 -> 13
    cases["phred33_read_as_64"] = _good(12) + (dict(qualmin=64),)
    # (an unterminated last line is in REF_UNDEFINED below)
    cases["no_trailing_newline"] = (f[:-1], r, {})
    cases["truncated_mid_line"] = (f[:len(f) // 2], r, {})
    cases["reverse_truncated_mid_line"] = (f, r[:len(r) // 3], {})
    for k in (1, 2, 3):
        cut = [i for i, c in enumerate(f) if c == 10][4 * 17 + k - 1] + 1
        cases[f"forward_ends_after_line_{k}"] = (f[:cut], r, {})
        cut = [i for i, c in enumerate(r) if c == 10][4 * 13 + k - 1] + 1
        cases[f"reverse_ends_after_line_{k}"] = (f, r[:cut], {})
    cases["reverse_ends_at_record"] = (f, r[:[i for i, c in enumerate(r) if c == 10][4 * 13 - 1] + 1], {})
    cases["empty"] = (b"", b"", {})
    lines = f.split(b"\n")
    cases["bad_nt"] = (_replace_line(f, 5, 1, lines[21][:30] + b"!" + lines[21][31:]), r, {})
    cases["bad_nt_reverse"] = (f, _replace_line(r, 7, 1, b"ACGT*" + r.split(b"\n")[29][5:]), {})
    cases["lowercase_and_iupac"] = (_replace_line(f, 3, 1, (b"acgtRYKMSWBDHVNXUn" * 9)[:150]), r, {})
    cases["missing_plus"] = (_replace_line(f, 4, 2, b"ACGT"), r, {})
    cases["junk_plus"] = (_replace_line(f, 4, 2, b"!junk"), r, {})
    cases["plus_with_text"] = (_replace_line(f, 4, 2, b"+M01271 repeated id"), r, {})
    cases["qual_short"] = (_replace_line(f, 6, 3, lines[27][:100]), r, {})
    cases["qual_long"] = (_replace_line(f, 6, 3, lines[27] + b"II"), r, {})
    cases["empty_forward_read"] = (_replace_line(_replace_line(f, 2, 1, b""), 2, 3, b""), r, {})
    cases["empty_reverse_read"] = (f, _replace_line(_replace_line(r, 2, 1, b""), 2, 3, b""), {})
    cases["empty_both_first"] = (_replace_line(_replace_line(f, 0, 1, b""), 0, 3, b""), _replace_line(_replace_line(r, 0, 1, b""), 0, 3, b""), {})
    cases["not_paired"] = (_replace_line(f, 9, 0, lines[0]), r, {})
    cases["same_direction"] = (f, _replace_line(r, 9, 0, lines[36]), {})
    cases["bad_id"] = (_replace_line(f, 11, 0, b"@garbage"), r, {})
    cases["bad_id_reverse"] = (f, _replace_line(r, 11, 0, b"@garbage"), {})
    cases["tag_absent_policy"] = (f, r, dict(policy=1))
    cases["tag_optional_policy"] = (f, r, dict(policy=2))
    cases["no_at_sign"] = (_replace_line(f, 1, 0, b"X" + lines[4][1:]), r, {})       # fastq.c:125 skips the first char unseen
    long_seq = (b"ACGT" * 120)[:450]
    cases["max_len"] = records([HEADERS[0]], [HEADERS[1]], [long_seq], [b"I" * 450], [long_seq[:100]], [b"5" * 100])
    cases["max_len"] += ({},)
    cases["too_long"] = records([HEADERS[0]], [HEADERS[1]], [long_seq + b"A"], [b"I" * 451], [long_seq[:100]], [b"5" * 100]) + ({},)
    cases["too_long_qual_450"] = records([HEADERS[0]], [HEADERS[1]], [long_seq + b"ACGT"], [b"I" * 450], [long_seq[:100]], [b"5" * 100]) + ({},)
    cases["high_quals"] = records([HEADERS[0]], [HEADERS[1]], [b"ACGT" * 5], [bytes(range(33, 53))], [b"ACGT" * 5], [bytes(range(107, 127))]) + ({},)
    cases["old_casava"] = records([HEADERS[11]] * 2, [HEADERS[12]] * 2, [b"ACGTACGTAC"] * 2, [b"IIIIIIIIII"] * 2, [b"GTACGTACGT"] * 2, [b"5555555555"] * 2) + (dict(policy=2),)
    cases["sra"] = records([HEADERS[19]], [HEADERS[19]], [b"ACGTACGTAC"], [b"IIIIIIIIII"], [b"GTACGTACGT"], [b"5555555555"]) + ({},)
    return cases


def rng_garbage(seed, n_lines=60):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTN@+:#/ 0123456789acgtIII\r!", dtype=np.uint8)
    out = []
    for _ in range(n_lines):
        ln = int(rng.integers(0, 40))
        out.append(bytes(rng.choice(alphabet, ln)))
    return b"\n".join(out) + b"\n"
