"""GPU parity of the lane-per-pair kernel (pb_lanes.cuh) and of its hand-over to the general kernel.

Configurations without per-base log p, primers or trims and with reads <= 256 nt are dispatched to the lane-per-pair
kernel; pairs it does not cover (N or IUPAC codes, qualities outside 0..46, no seed, tiny reads) are appended to a
deferral list and assembled by the general kernel in a second launch.  Either way the result must equal the oracle's:
integer fields and merged bases bit-exact, quality / overlap score within 1e-6 (the lane kernel adds in the reference's
order, so its quality is in fact identical)."""
import numpy as np
import pytest

import datasets
import oracle_lib
import pandaseq_b200 as pb
from pandaseq_b200 import synth
from parity import compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = pb.Context(0)
    yield c
    c.close()


def run_lanes(ctx, cfg, batch, *, expect_lanes=True, seq_stride=None):
    before = ctx.lanes_stats()
    got = ctx.assemble_host(cfg, batch, want_nt=True, want_p=False, seq_stride=seq_stride)
    after = ctx.lanes_stats()
    want = oracle_lib.assemble("port", cfg, batch)
    rep = compare(got, want)
    rep["lanes_pairs"] = after[0] - before[0]
    rep["deferred"] = after[1] - before[1]
    if expect_lanes is True:
        assert rep["lanes_pairs"] == batch.n, rep
    elif expect_lanes is False:
        assert rep["lanes_pairs"] == 0, rep
    return got, want, rep


def clean(n, seed=1, tmpl=(180, 280), rl=(150, 150)):
    """A/C/G/T only, qualities 2..41: what the lane-per-pair kernel keeps for itself."""
    return synth.generate(n, rl=rl, tmpl=tmpl, seed=seed).to_flat()


@pytest.mark.parametrize("algo", ["simple_bayesian", "uparse", "flash"])
def test_clean_pairs_stay_on_the_lane_kernel(ctx, algo):
    b = clean(20_000, seed=101)
    got, want, rep = run_lanes(ctx, pb.make_config(algo), b)
    assert rep["ok"], rep
    assert rep["deferred"] <= b.n // 500, rep         # only the pairs without any seed (SLOW) are handed on
    assert rep["max_dq"] <= 1e-12 and rep["max_dp"] == 0.0, rep


@pytest.mark.parametrize("algo", ["simple_bayesian", "uparse", "flash"])
def test_cfg1_with_n_and_b_tails(ctx, algo):
    # 0.1 % N and 5 % '#' tails: about a quarter of the pairs carry an N and are handed to the general kernel
    got, want, rep = run_lanes(ctx, pb.make_config(algo), datasets.cfg1())
    assert rep["ok"], rep
    assert 0 < rep["deferred"] < rep["n"], rep


@pytest.mark.parametrize("maxoverlap", [0, 140, 300])
def test_read_through_stress(ctx, maxoverlap):
    # inserts shorter than the reads: no seed inside the bitset (SLOW), overlaps longer than a read when maxoverlap allows
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian", maxoverlap=maxoverlap), datasets.stress())
    assert rep["ok"], rep


def test_b_tails_without_n(ctx):
    b = synth.generate(8000, rl=(150, 150), tmpl=(160, 280), seed=31, btail_rate=0.5).to_flat()
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), b)
    assert rep["ok"], rep
    assert rep["deferred"] <= b.n // 4, rep           # short overlaps under a long low-quality tail have no seed
    assert rep["max_dq"] <= 1e-12, rep


def test_short_and_unequal_reads(ctx):
    rng = np.random.default_rng(5)
    pairs = []
    for i in range(3000):
        F, R = int(rng.integers(16, 161)), int(rng.integers(16, 161))
        L = int(rng.integers(max(F, R), F + R - 1))
        t = rng.integers(0, 4, size=L)
        f, r = t[:F].copy(), t[::-1][:R].copy()
        for arr in (f, r):
            m = rng.random(len(arr)) < 0.02
            arr[m] = (arr[m] + rng.integers(1, 4, size=int(m.sum()))) & 3
        pairs.append((1 << f, rng.integers(0, 47, size=F), 1 << r, rng.integers(0, 47, size=R)))
    b = synth.FlatBatch.from_pairs(pairs)
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), b)
    assert rep["ok"], rep
    assert rep["max_dq"] <= 1e-12, rep


def test_thresholds_minoverlap_and_error_estimate(ctx):
    b = clean(6000, seed=9, tmpl=(152, 296))
    for kw in (dict(threshold=0.9), dict(threshold=0.3, minoverlap=30), dict(minoverlap=140, maxoverlap=145), dict(sb_q=0.1),
               dict(maxoverlap=100), dict(minoverlap=149)):
        got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian", **kw), b)
        assert rep["ok"], (kw, rep)


def test_low_complexity(ctx):
    # homopolymers and short tandem repeats: the same code at many positions, long probe chains, many candidates
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), datasets.low_complexity())
    assert rep["ok"], rep


def test_edge_cases_mix_of_classes(ctx):
    # reads of every size next to each other: the chunk holding the 450-nt reads goes to the general kernel as a whole
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), datasets.edge_cases(), expect_lanes=None)
    assert rep["ok"], rep


def test_edge_cases_up_to_160(ctx):
    full = datasets.edge_cases()
    fl, rl = full.lengths()
    keep = [i for i in range(full.n) if fl[i] <= 160 and rl[i] <= 160]
    pairs = []
    for i in keep:
        f, r = full.pair(i)
        pairs.append((f[:, 0], f[:, 1], r[:, 0], r[:, 1]))
    b = synth.FlatBatch.from_pairs(pairs)
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), b)
    assert rep["ok"], rep
    got, want, rep = run_lanes(ctx, pb.make_config("flash"), b)
    assert rep["ok"], rep


@pytest.mark.parametrize("filters", [f for f in datasets.FILTER_SETS if not any(k == "min_phred" for k, _ in f)])
def test_filters_on_the_result_record(ctx, filters):
    b = datasets.cfg1(4000)
    cfg = pb.make_config("simple_bayesian", filters=filters)
    got, want, rep = run_lanes(ctx, cfg, b)
    assert rep["ok"], rep


def test_row_capacity_smaller_than_the_sequence(ctx):
    b = clean(2000, seed=12)
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), b, seq_stride=208)
    res = got["results"]
    assert (res["status"] == want["status"]).all() and (res["seq_len"] == want["seq_len"]).all()
    w = 208
    sl = np.minimum(want["seq_len"], w)
    ok = want["status"] == 0
    mask = (np.arange(w)[None, :] < sl[:, None]) & ok[:, None]
    assert not ((got["seq_nt"][:, :w] != want["seq_nt"][:, :w]) & mask).any()


def test_trims_and_per_base_p_use_the_general_kernel(ctx):
    b = clean(1000, seed=13)
    run_lanes(ctx, pb.make_config("simple_bayesian", forward_trim=5), b, expect_lanes=False)
    before = ctx.lanes_stats()
    ctx.assemble_host(pb.make_config("simple_bayesian"), b, want_nt=True, want_p=True)
    assert ctx.lanes_stats()[0] == before[0]
    run_lanes(ctx, pb.make_config("rdp_mle"), b, expect_lanes=False)


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "flash", "uparse"])
def test_reads_up_to_256_nt(ctx, algo):
    # the 256-nt class of the seeding kernel (8 mask words) and of the lane kernel (784-byte record slots, 7 warps)
    got, want, rep = run_lanes(ctx, pb.make_config(algo), datasets.long250(3000))
    assert rep["ok"], rep
    clean250 = synth.generate(6000, rl=(250, 250), tmpl=(260, 480), seed=77).to_flat()
    got, want, rep = run_lanes(ctx, pb.make_config(algo), clean250)
    assert rep["ok"], rep
    assert rep["deferred"] <= clean250.n // 100, rep


def test_pear_on_the_lane_kernel(ctx):
    # algo_pear.c:32-59 term by term in the lane kernel: the overlap score is the reference's sum in the reference's order
    for batch in (clean(20_000, seed=5), datasets.cfg1(), datasets.stress(), datasets.low_complexity()):
        got, want, rep = run_lanes(ctx, pb.make_config("pear"), batch)
        assert rep["ok"], rep
    got, want, rep = run_lanes(ctx, pb.make_config("pear"), clean(20_000, seed=6))
    ok = want["status"] == 0
    # ... for the pairs the lane kernel keeps; the few it hands on get the general kernel's shuffle-tree sum (last bits)
    differ = got["results"]["est_prob"][ok].view(np.uint64) != want["est_prob"][ok].view(np.uint64)
    assert int(differ.sum()) <= rep["deferred"] and rep["deferred"] <= 40, rep
    assert np.abs(got["results"]["est_prob"][ok] - want["est_prob"][ok]).max() <= 1e-9
    # reverse longer than forward: algo_pear.c:52 reads past the forward read; those pairs are handed to the general kernel
    rng = np.random.default_rng(8)
    pairs = []
    for i in range(2000):
        F, R = int(rng.integers(40, 200)), int(rng.integers(40, 200))
        L = int(rng.integers(max(F, R), F + R - 8))
        t = rng.integers(0, 4, size=L)
        pairs.append((1 << t[:F], rng.integers(2, 42, size=F), 1 << t[::-1][:R], rng.integers(2, 42, size=R)))
    b = synth.FlatBatch.from_pairs(pairs)
    got, want, rep = run_lanes(ctx, pb.make_config("pear", minoverlap=8), b)
    assert rep["ok"], rep
    assert rep["deferred"] > 300, rep


def test_pear_with_qualities_above_the_clamp(ctx):
    # qualities above 46 (PHREDCLAMP, prob.h:23): the lane kernel scores pear's candidates from the low six bits of the raw byte, so
    # a verdict reached that way (NOALGN included) must not stand -- the pair has to reach the general kernel, which clamps
    rng = np.random.default_rng(19)
    pairs = []
    for i in range(6000):
        F, R = int(rng.integers(90, 151)), int(rng.integers(60, 91))
        L = int(rng.integers(F, F + R - 20))
        t = rng.integers(0, 4, size=L)
        f, r = t[:F].copy(), t[::-1][:R].copy()
        # heavy damage in the overlap so that many pairs sit near the no-alignment verdict
        hit = rng.random(R) < rng.choice([0.0, 0.3, 0.5, 0.7])
        r[hit] = (r[hit] + rng.integers(1, 4, size=int(hit.sum()))) % 4
        hi = 94 if i % 3 else 64
        pairs.append((1 << f, rng.integers(30, hi, size=F), 1 << r, rng.integers(30, hi, size=R)))
    b = synth.FlatBatch.from_pairs(pairs)
    got, want, rep = run_lanes(ctx, pb.make_config("pear", minoverlap=8), b)
    assert rep["ok"], rep
    assert int((want["status"] == 4).sum()) > 50 and int((want["status"] == 0).sum()) > 500, np.bincount(want["status"])


def test_mixed_lengths_up_to_256_nt(ctx):
    b = synth.generate(5000, rl=(75, 256), tmpl=None, seed=15, mixed=True, n_rate=0.0005, btail_rate=0.05).to_flat()
    for algo in ("simple_bayesian", "pear"):
        got, want, rep = run_lanes(ctx, pb.make_config(algo), b)
        assert rep["ok"], (algo, rep)


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "flash", "uparse"])
def test_length_classes_up_to_320_nt(ctx, algo):
    # mixed batches are listed by length class (<= 160, <= 256, <= 320 nt: pb::class_list_kernel) and every class runs the sweep and
    # the lane kernel sized for it; BASELINE config 5's shape (2x(75-300)), with N and '#' tails so that pairs are handed on as well
    b = synth.generate(6000, rl=(75, 300), tmpl=None, seed=25, mixed=True, n_rate=0.0005, btail_rate=0.05).to_flat()
    got, want, rep = run_lanes(ctx, pb.make_config(algo), b)
    assert rep["ok"], (algo, rep)
    clean = synth.generate(6000, rl=(75, 300), tmpl=None, seed=26, mixed=True).to_flat()
    got, want, rep = run_lanes(ctx, pb.make_config(algo), clean)
    assert rep["ok"], (algo, rep)
    if algo != "pear":          # pear hands on every pair whose reverse read is the longer one (algo_pear.c:52 reads past the forward read)
        assert rep["deferred"] <= clean.n // 50, rep
    # 2x300 exactly: everything in the 320-nt class
    b300 = synth.generate(3000, rl=(300, 300), tmpl=(320, 580), seed=27, n_rate=0.0002).to_flat()
    got, want, rep = run_lanes(ctx, pb.make_config(algo), b300)
    assert rep["ok"], (algo, rep)


def test_length_classes_with_reads_up_to_450_nt(ctx):
    # pairs with a read above 320 nt go straight to the general kernel's list; the rest of the batch stays on the two-kernel path
    rng = np.random.default_rng(31)
    pairs = []
    for i in range(3000):
        F, R = int(rng.integers(20, 451)), int(rng.integers(20, 451))
        L = int(rng.integers(max(F, R), F + R - 12))
        t = rng.integers(0, 4, size=L)
        f, r = t[:F].copy(), t[::-1][:R].copy()
        for arr in (f, r):
            m = rng.random(len(arr)) < 0.01
            arr[m] = (arr[m] + rng.integers(1, 4, size=int(m.sum()))) & 3
        pairs.append((1 << f, rng.integers(2, 42, size=F), 1 << r, rng.integers(2, 42, size=R)))
    b = synth.FlatBatch.from_pairs(pairs)
    for algo in ("simple_bayesian", "pear"):
        got, want, rep = run_lanes(ctx, pb.make_config(algo, minoverlap=10), b)
        assert rep["ok"], (algo, rep)
        assert 500 < rep["deferred"] < b.n, rep


def test_length_classes_small_batches_filters_and_no_rows(ctx):
    # the length-class path at its edges: one pair, one more than a warp-batch, classes that stay empty, checks on the result record,
    # a caller that does not want the merged reads, an explicit maxoverlap (hash-join seeding: whole batches, reads up to 256 nt only)
    big = synth.generate(2000, rl=(75, 300), tmpl=None, seed=41, mixed=True, n_rate=0.0003).to_flat()
    for n in (1, 2, 33, 65):
        b = big.slice(0, n)
        got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian"), b)
        assert rep["ok"], (n, rep)
    only300 = synth.generate(40, rl=(300, 300), tmpl=(330, 560), seed=42).to_flat()          # classes 0 and 1 stay empty
    got, want, rep = run_lanes(ctx, pb.make_config("flash"), only300)
    assert rep["ok"], rep
    for filters in ([("short", 250), ("long", 420)], [("pear_test", (1.0, -1.0, 0.01))], [("min_overlapbits", 80.0), ("completely_miss_the_point", 2)]):
        got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian", filters=filters), big)
        assert rep["ok"], (filters, rep)
    before = ctx.lanes_stats()
    got = ctx.assemble_host(pb.make_config("simple_bayesian"), big, want_nt=False)
    assert ctx.lanes_stats()[0] - before[0] == big.n
    want = oracle_lib.assemble("port", pb.make_config("simple_bayesian"), big)
    rep = compare(got, want)
    assert rep["ok"], rep
    # reads above 256 nt with an explicit maxoverlap: no sweep, no hash-join class for them -> the general kernel takes the batch
    got, want, rep = run_lanes(ctx, pb.make_config("simple_bayesian", maxoverlap=200), big, expect_lanes=False)
    assert rep["ok"], rep
