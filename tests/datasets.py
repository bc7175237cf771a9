"""Seeded input sets for the parity tests (small enough for the CPU oracle to finish in seconds)."""
from __future__ import annotations

import numpy as np

from pandaseq_b200 import synth


def cfg1(n=10_000):
    """BASELINE config 1 shape (2x150, insert 180-280) with N and B-tail decoration."""
    return synth.generate_config(1, n=n, n_rate=0.001, btail_rate=0.05).to_flat()


def stress(n=4000, seed=77):
    """Read-through inserts (40-200 nt under 2x150 reads), 0.5 % N, 20 % '#' tails (SURVEY.md §8a probe set)."""
    return synth.generate(n, rl=(150, 150), tmpl=(40, 200), seed=seed, n_rate=0.005, btail_rate=0.2).to_flat()


def mixed(n=4000):
    """BASELINE config 5 shape: independent read lengths 75-300."""
    return synth.generate_config(5, n=n, n_rate=0.001, btail_rate=0.05).to_flat()


def long250(n=3000):
    return synth.generate_config(3, n=n, n_rate=0.001, btail_rate=0.05).to_flat()


def primers300(n=2000):
    return synth.generate_config(4, n=n, n_rate=0.001, btail_rate=0.02).to_flat()


def primer_codes():
    fwd = synth.encode(synth.FWD_PRIMER)
    # the assembler stores the reverse primer complemented (args_assembler.c:222)
    rev = synth.encode("".join(synth._COMP[c] for c in synth.REV_PRIMER))
    return fwd, rev


def low_complexity(n=600, seed=5):
    """Homopolymer / short tandem repeat reads: many identical 8-mers, i.e. lost k-mers (FML) and long hash chains."""
    rng = np.random.default_rng(seed)
    pairs = []
    for i in range(n):
        L = int(rng.integers(160, 280))
        kind = i % 4
        if kind == 0:
            t = np.full(L, rng.integers(0, 4))
        elif kind == 1:
            unit = rng.integers(0, 4, size=int(rng.integers(2, 6)))
            t = np.tile(unit, L // len(unit) + 1)[:L]
        elif kind == 2:
            t = rng.integers(0, 4, size=L)
            a = int(rng.integers(0, L - 60))
            t[a:a + 60] = t[a]
        else:
            unit = rng.integers(0, 4, size=int(rng.integers(9, 20)))
            t = np.tile(unit, L // len(unit) + 1)[:L]
        # sprinkle a few errors
        f = t[:150].copy()
        r = t[::-1][:150].copy()          # template order reversed = reverse read (complemented, read order)
        for arr in (f, r):
            m = rng.random(150) < 0.01
            arr[m] = (arr[m] + rng.integers(1, 4, size=int(m.sum()))) & 3
        fq = rng.integers(2, 42, size=150)
        rq = rng.integers(2, 42, size=150)
        pairs.append((1 << f, fq, 1 << r, rq))
    return synth.FlatBatch.from_pairs(pairs)


def edge_cases():
    """Empty, 1-base, tiny, unequal, all-N, degenerate codes, out-of-range qualities, maximum length."""
    rng = np.random.default_rng(11)

    def rnd(L):
        return 1 << rng.integers(0, 4, size=L)

    def q(L, lo=2, hi=42):
        return rng.integers(lo, hi, size=L)

    pairs = []
    e = np.zeros(0, dtype=np.int64)
    pairs.append((e, e, e, e))                                   # empty / empty
    pairs.append((rnd(1), q(1), rnd(1), q(1)))                   # 1 / 1
    pairs.append((rnd(150), q(150), e, e))                       # 150 / empty
    pairs.append((rnd(2), q(2), rnd(2), q(2)))                   # 2 / 2
    pairs.append((rnd(5), q(5), rnd(300), q(300)))               # tiny / long
    pairs.append((rnd(300), q(300), rnd(9), q(9)))
    t = rng.integers(0, 4, size=20)
    pairs.append((1 << t[:12], q(12), 1 << t[::-1][:12], q(12)))  # short but overlapping, < 9 bases: no k-mers at all
    t = rng.integers(0, 4, size=620)
    pairs.append((1 << t[:450], q(450), 1 << t[::-1][:450], q(450)))   # maximum length both
    t = rng.integers(0, 4, size=450)
    pairs.append((1 << t[:450], q(450), 1 << t[::-1][:450], q(450)))   # complete overlap at maximum length
    pairs.append((np.full(150, 15), np.full(150, 2), np.full(150, 15), np.full(150, 2)))   # all N, all '#'
    t = rng.integers(0, 4, size=200)
    f, r = 1 << t[:150], 1 << t[::-1][:150]
    f2 = f.copy(); f2[60:150:7] = 15                              # N every 7 bases: no valid k-mer in the overlap
    pairs.append((f2, q(150), r, q(150)))
    f3 = f.copy(); f3[100:140:3] |= rnd(14)                       # degenerate IUPAC codes
    r3 = r.copy(); r3[10:40:4] = 0                                # invalid code 0
    pairs.append((f3, q(150), r3, q(150)))
    pairs.append((f, q(150, -128, 128), r, q(150, -128, 128)))    # every char value as quality
    pairs.append((f, np.full(150, 2), r, np.full(150, 2)))        # fully B-masked
    pairs.append((f, np.full(150, 46), r, np.full(150, 47)))      # PHREDMAX boundary
    pairs.append((f, np.full(150, 0), r, np.full(150, 0)))        # PHRED 0 (score -2 special case)
    for L in (8, 9, 10, 16, 17, 31, 32, 33, 63, 64, 65):          # word-boundary lengths
        t = rng.integers(0, 4, size=L + L // 2)
        pairs.append((1 << t[:L], q(L), 1 << t[::-1][:L], q(L)))
    return synth.FlatBatch.from_pairs(pairs)


def overhang(n=2000, seed=21):
    """Amplicons shorter than the reads (inserts 60-200 nt between the two 17-nt primers, 2x150 reads): most reads run
    through the far primer into unrelated sequence -- what the overhang trimmer (hang.c) cuts off."""
    return synth.generate(n, rl=(150, 150), tmpl=(60, 200), seed=seed, primers=True, n_rate=0.002, btail_rate=0.05).to_flat()


def overhang_codes():
    """(-P, -Q) sequences as panda_trim_overhangs receives them (args_hang.c:105-107): what follows the insert in the forward
    read, and -- complemented at parse time -- what follows it in the reverse read."""
    fwd = synth.encode(synth.revcomp(synth.REV_PRIMER))
    rev = synth.encode("".join(synth._COMP[c] for c in synth.revcomp(synth.FWD_PRIMER)))
    return fwd, rev


# pear_test with other parameters: checked against the oracle restatement only (oracle/ref_harness.c says why)
PEAR_TEST_PARAMS = [(1.0, -1.0, 1e-9), (1.0, -1.0, 0.9), (2.0, -3.0, 0.01), (1.0, -0.5, 1e-6), (0.5, 0.5, 0.01), (1.0, 2.0, 0.3)]

FILTER_SETS = [
    [("no_n", 0)],
    [("short", 200), ("long", 240)],
    [("long", 230), ("no_n", 0), ("short", 190)],
    [("min_overlapbits", 120.0)],
    [("completely_miss_the_point", 1)],
    [("min_phred", 12)],
    [("pear_test", (1.0, -1.0, 0.01))],                       # the plugin's defaults: the only values the reference can run it with
    [("short", 185), ("pear_test", (1.0, -1.0, 0.01)), ("long", 270)],
    [("min_phred", 4), ("completely_miss_the_point", 3), ("min_overlapbits", 100.5), ("no_n", 0), ("short", 185), ("long", 270)],
]
