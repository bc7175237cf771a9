"""The primer scan's pruning rule (pb_kernels.cuh, primer_offset, DESIGN.md section 5 "primer scan"), restated in Python and checked on the
CPU: a start offset is summed exactly only if the bound  m * emax / (index + 1)  (m = bases among the primer's first eight that disagree,
emax = qual_score_err at the read's lowest quality) still reaches the best value so far.  The pruned scan must return what the full scan
returns -- the reference's offset.c:47-112, here both as a plain log-domain loop and as the oracle's compiled port -- on reads with and
without the primer, with errors in it, with low, zero, out-of-range and negative qualities, N / IUPAC codes on either side, repeats
(several equally good starts), and primers from 1 to 40 bases."""
import numpy as np
import pytest

import oracle_lib

T = oracle_lib.tables("port")
SCORE, SCORE_ERR = np.asarray(T["score"], dtype=np.float64), np.asarray(T["score_err"], dtype=np.float64)


def clampq(q):
    q = np.asarray(q, dtype=np.int64)
    q = np.where(q >= 128, q - 256, q)          # the reference's qualities are (signed) chars
    return np.clip(q, 0, 46)                    # PHREDCLAMP, prob.h:23


def terms(seq_nt, q, primer, s):
    """the addends of start s in primer order: (value, disagrees) for every primer base that is not N (offset.c:93-101)"""
    out = []
    for x, pn in enumerate(primer):
        if pn == 15:
            continue
        hit = (int(seq_nt[s + x]) & pn) != 0
        out.append(SCORE[q[s + x]] if hit else SCORE_ERR[q[s + x]])
    return out


def exact(seq_nt, q, primer, s):
    total = 0.0
    for v in terms(seq_nt, q, primer, s):
        total += v
    return total / float(s + len(primer) + 1)


def full_scan(seq_nt, q, primer, threshold_log):
    P, n = len(primer), len(seq_nt)
    if P > n:
        return 0
    best, best_index = P * threshold_log, 0
    for s in range(n - P):                     # the start ending exactly at the read's end is never examined
        v = exact(seq_nt, q, primer, s)
        if v > best:                           # strictly: the first of equals stays
            best, best_index = v, s + P + 1
    return best_index


def pruned_scan(seq_nt, q, primer, threshold_log, stats):
    P, n = len(primer), len(seq_nt)
    if P > n:
        return 0
    nstart = n - P
    if nstart <= 0:
        return 0
    best, best_s = P * threshold_log, -1
    emax = SCORE_ERR[q.min()]
    first = [x for x in range(min(8, P)) if primer[x] != 15]
    m = np.array([sum((int(seq_nt[s + x]) & primer[x]) == 0 for x in first) for s in range(nstart)])
    s0 = int(np.lexsort((np.arange(nstart), m))[0])
    v0 = exact(seq_nt, q, primer, s0)
    stats["exact"] += 1
    if v0 > best:
        best, best_s = v0, s0
    clean = int((m == 0).sum())
    if emax * 0.999999999 < best * float(n) and clean == (1 if m[s0] == 0 else 0):
        stats["short"] += 1
        return 0 if best_s < 0 else best_s + P + 1
    for base in range(0, nstart, 32):
        cands = [s for s in range(base, min(base + 32, nstart))
                 if s != s0 and not (float(m[s]) * emax * 0.999999999 < best * float(s + P + 1))]
        if not cands:
            continue
        stats["exact"] += len(cands)
        vals = [(exact(seq_nt, q, primer, s), -s) for s in cands]
        v, negs = max(vals)
        who = -negs
        if v > best or (v == best and best_s >= 0 and who < best_s):
            best, best_s = v, who
    return 0 if best_s < 0 else best_s + P + 1


def make_case(rng, kind):
    n = int(rng.integers(12, 301))
    P = int(rng.integers(1, min(40, n) + 1))
    nt = (1 << rng.integers(0, 4, size=n)).astype(np.uint8)
    primer = (1 << rng.integers(0, 4, size=P)).astype(np.uint8)
    qual = rng.integers(2, 42, size=n).astype(np.uint8)
    if kind % 2 == 0 and n - P > 1:             # plant the primer, with a few errors now and then
        at = int(rng.integers(0, n - P))
        nt[at:at + P] = primer
        for _ in range(int(rng.integers(0, 3))):
            k = at + int(rng.integers(0, P))
            nt[k] = 1 << ((int(np.log2(nt[k])) + 1 + int(rng.integers(0, 3))) & 3)
    if kind % 3 == 0:                           # IUPAC / N in the primer and the read
        for arr in (primer, nt):
            sel = rng.random(len(arr)) < 0.08
            arr[sel] = rng.integers(1, 16, size=int(sel.sum()))
    if kind % 5 == 0:                           # low, zero, out-of-range and "negative" qualities
        sel = rng.random(n) < 0.15
        qual[sel] = rng.choice([0, 1, 2, 2, 2, 47, 60, 93, 130, 255], size=int(sel.sum()))
    if kind % 7 == 0 and n > 3 * P:             # a repeat: the same stretch twice, several equally good starts
        nt[P:2 * P] = nt[0:P]
        qual[P:2 * P] = qual[0:P]
    if kind % 11 == 0:                          # one quality everywhere: many exact ties between starts
        qual[:] = 30
    return nt, qual, primer


@pytest.mark.parametrize("reverse", [False, True])
def test_pruned_scan_equals_the_full_scan(reverse):
    rng = np.random.default_rng(7 + reverse)
    stats = {"exact": 0, "short": 0}
    starts = 0
    for kind in range(900):
        nt, qual, primer = make_case(rng, kind)
        threshold_log = float(np.log(rng.choice([0.6, 0.9, 0.3])))
        read = np.stack([nt, qual], axis=1).astype(np.uint8)
        seq = read[::-1] if reverse else read     # offset.c:79: a reverse scan starts at the read's end
        q = clampq(seq[:, 1])
        want = full_scan(seq[:, 0], q, [int(p) for p in primer], threshold_log)
        got = pruned_scan(seq[:, 0], q, [int(p) for p in primer], threshold_log, stats)
        assert got == want, (kind, reverse, len(nt), len(primer))
        ref = oracle_lib.compute_offset("port", threshold_log, 0.0, reverse, read, bytes(primer))
        assert ref == want, (kind, reverse, "the log-domain scan differs from the oracle's")
        starts += max(len(nt) - len(primer), 0)
    assert stats["exact"] <= starts + 900


def test_reads_that_carry_their_primer_are_settled_after_one_sum():
    """BASELINE config 4's shape: the primer once, without error, qualities 20..41 -- the bound removes every other start.  (A read
    with a few very low qualities anywhere weakens the bound for all of its starts, emax being taken over the whole read: such reads
    fall back to summing most starts, as the first test's uniform 2..41 qualities show.)"""
    rng = np.random.default_rng(3)
    stats = {"exact": 0, "short": 0}
    cases = 0
    for _ in range(300):
        n, P = int(rng.integers(100, 301)), int(rng.integers(12, 26))
        nt = (1 << rng.integers(0, 4, size=n)).astype(np.uint8)
        primer = (1 << rng.integers(0, 4, size=P)).astype(np.uint8)
        at = int(rng.integers(0, 20))
        nt[at:at + P] = primer
        qual = rng.integers(20, 42, size=n)
        q = clampq(qual)
        want = full_scan(nt, q, [int(p) for p in primer], float(np.log(0.6)))
        got = pruned_scan(nt, q, [int(p) for p in primer], float(np.log(0.6)), stats)
        assert got == want == at + P + 1
        cases += 1
    assert stats["short"] >= cases - 3 and stats["exact"] <= cases + 30, (stats, cases)


def test_what_the_bound_rests_on():
    # every addend is a log-probability (<= 0), and qual_score_err falls with the quality: emax is its value at the read's lowest quality
    assert (SCORE[:47] <= 0).all() and (SCORE_ERR[:47] <= 0).all()
    assert (np.diff(SCORE_ERR[:47]) <= 0).all()
