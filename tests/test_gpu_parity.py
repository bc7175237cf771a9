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


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle", "flash", "ea_util", "stitch", "uparse"])
def test_cfg1_all_algorithms(ctx, algo):
    got, want, rep = run_both(ctx, pb.make_config(algo), datasets.cfg1())
    assert rep["ok"], rep
    assert rep["max_dq"] <= 1e-9 and rep["max_dp"] <= 1e-9


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle", "ea_util", "stitch", "uparse"])
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


@pytest.mark.parametrize("plen", [1, 7, 8, 9, 15, 16, 17, 24, 31])
def test_primers_with_n_and_iupac_codes(ctx, plen):
    """Primer lengths either side of the 8-base blocks of the scan; N (contributes nothing, offset.c:97), IUPAC codes and the
    invalid code 0 inside the primer; with and without a penalty; mixed read lengths (a primer longer than the read)."""
    rng = np.random.default_rng(100 + plen)
    fwd = 1 << rng.integers(0, 4, size=plen)
    rev = 1 << rng.integers(0, 4, size=plen)
    for cfg_kw in (dict(), dict(primer_penalty=0.001)):
        for variant in range(3):
            f, r = fwd.copy(), rev.copy()
            if variant >= 1:
                f[rng.integers(0, plen)] = 15
                r[rng.integers(0, plen)] |= 1 << rng.integers(0, 4)
            if variant == 2:
                r[rng.integers(0, plen)] = 15
                f[rng.integers(0, plen)] = 0
            cfg = pb.make_config("simple_bayesian", forward_primer=f, reverse_primer=r, **cfg_kw)
            for batch in (datasets.cfg1(600), datasets.mixed(400)):
                got, want, rep = run_both(ctx, cfg, batch)
                assert rep["ok"], (plen, variant, cfg_kw, rep)


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
    for algo in ("simple_bayesian", "pear", "rdp_mle", "flash", "ea_util", "stitch", "uparse"):
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
        ctx.assemble_host(pb.make_config(9), b)


def test_pinned_host_buffers_take_the_direct_copy_path(ctx):
    """pb_assemble_host copies straight from/to page-locked caller buffers (two streams); same results as staging."""
    import ctypes as C
    import torch
    b = datasets.cfg1(5000)
    cfg = pb.make_config("simple_bayesian")
    want = ctx.assemble_host(cfg, b, want_nt=True, want_p=False)       # pageable numpy -> staged path
    stride = want["seq_stride"]

    def pin(a, dt=None):
        t = torch.empty(a.shape, dtype=dt or torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    f, fo = pin(b.f_data), pin(b.f_off.view(np.int64))
    r, ro = pin(b.r_data), pin(b.r_off.view(np.int64))
    res = torch.zeros((b.n, 32), dtype=torch.uint8, pin_memory=True)
    nt = torch.zeros((b.n, stride // 2), dtype=torch.uint8, pin_memory=True)
    cnt = np.zeros(pb.PB_NCOUNTERS, np.int64)
    rc = pb.lib().pb_assemble_host(ctx._h, C.byref(cfg), b.n, f.data_ptr(), fo.data_ptr(), r.data_ptr(), ro.data_ptr(),
                                   res.data_ptr(), nt.data_ptr(), None, stride, cnt.ctypes.data)
    assert rc == 0, pb.lib().pb_last_error()
    assert np.array_equal(res.numpy().view(pb.PAIR_RESULT_DTYPE).ravel(), want["results"])
    assert np.array_equal(nt.numpy(), want["seq_nt_packed"])
    assert np.array_equal(cnt, want["counters"])


def test_device_resident_entry_point(ctx):
    """pb_pack_device + pb_assemble_device on torch-owned HBM (what bench.py times) == the host-buffer path."""
    import torch
    from pandaseq_b200 import synth
    rect = synth.generate_config(1, n=3000, device="cuda", n_rate=0.001, btail_rate=0.05)
    f_data, f_off, r_data, r_off = rect.to_flat_tensors()
    reads, meta, max_len, total = ctx.pack_device(f_data, f_off, r_data, r_off)
    n = f_off.numel() - 1
    stride = (2 * max_len + 15) & ~15
    res = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    nt = torch.zeros((n, stride // 2), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64, device="cuda")
    cfg = pb.make_config("simple_bayesian")
    torch.cuda.synchronize()
    ctx.assemble_device(cfg, n, max_len, reads, meta, res, nt, None, stride, cnt)
    ctx.synchronize()
    flat = synth.FlatBatch(f_data.cpu().numpy(), f_off.cpu().numpy().astype(np.uint64), r_data.cpu().numpy(), r_off.cpu().numpy().astype(np.uint64))
    got = dict(results=res.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel(), seq_nt=pb.unpack_nt(nt.cpu().numpy()), seq_p=None,
               counters=cnt.cpu().numpy())
    rep = compare(got, oracle_lib.assemble("port", cfg, flat))
    assert rep["ok"], rep


@pytest.mark.parametrize("algo", ["simple_bayesian", "rdp_mle", "pear"])
def test_primers_after_assembly(ctx, algo):
    """-a / primers_after: the primers are located on the assembled sequence (assembler.c:300-333, offset.c:114-133)."""
    fwd, rev = datasets.primer_codes()
    cases = [(datasets.primers300(1200), dict(forward_primer=fwd, reverse_primer=rev)),
             (datasets.primers300(600), dict(forward_primer=fwd, reverse_primer=rev, primer_penalty=0.0005)),
             (datasets.primers300(600), dict(forward_primer=fwd, reverse_trim=9)),
             (datasets.cfg1(1500), dict(forward_trim=12, reverse_trim=7)),
             (datasets.cfg1(800), dict(forward_primer=fwd, reverse_primer=rev)),            # reads without the primers: all NOFP
             (datasets.stress(1000), dict(forward_primer=fwd[:6], reverse_primer=rev[:5], maxoverlap=300)),
             (datasets.edge_cases(), dict(forward_primer=fwd[:3], reverse_trim=2))]
    for batch, kw in cases:
        got, want, rep = run_both(ctx, pb.make_config(algo, post_primers=True, **kw), batch)
        assert rep["ok"], (kw.keys(), rep)


# ---- the stages SURVEY.md 8f ranks 3 and 4: overhang trimmer (hang.c) and filters on the assembled pair (module.c) ----
@pytest.mark.parametrize("filters", datasets.FILTER_SETS)
@pytest.mark.parametrize("algo", ["simple_bayesian", "rdp_mle"])
def test_filters(ctx, filters, algo):
    got, want, rep = run_both(ctx, pb.make_config(algo, filters=filters), datasets.cfg1(3000))
    assert rep["ok"], rep
    assert (got["results"]["status"] >= 8).sum() == want["counters"][pb.C_REJECTED:pb.C_REJECTED + 7].sum()


@pytest.mark.parametrize("params", datasets.PEAR_TEST_PARAMS)
def test_pear_test_other_parameters(ctx, params):
    """plugin_pear_test.c with parameters the reference's own argument parsing cannot deliver (oracle/ref_harness.c), and on
    pairs whose limit goes negative (where the plugin's loop never ends; defined as the sum of all terms): device vs oracle,
    general kernel (per-base p requested) and two-kernel path, read-through and mixed-length sets, three scorers."""
    for algo, batch in (("simple_bayesian", datasets.stress(1500)), ("pear", datasets.mixed(1000)), ("flash", datasets.cfg1(800)),
                        ("rdp_mle", datasets.overhang(800))):
        cfg = pb.make_config(algo, filters=[("pear_test", params), ("short", 60)])
        got, want, rep = run_both(ctx, cfg, batch)
        assert rep["ok"], (algo, params, rep)
        got = ctx.assemble_host(cfg, batch, want_nt=True, want_p=False)
        rep = compare(got, want)
        assert rep["ok"], (algo, params, "no per-base p", rep)


def test_filters_after_primer_strip(ctx):
    fwd, rev = datasets.primer_codes()
    cfg = pb.make_config("simple_bayesian", forward_primer=fwd, reverse_primer=rev, post_primers=True,
                         filters=[("short", 440), ("min_phred", 3), ("long", 480)])
    got, want, rep = run_both(ctx, cfg, datasets.primers300(600))
    assert rep["ok"], rep


@pytest.mark.parametrize("skip", [False, True])
@pytest.mark.parametrize("which", ["both", "forward", "reverse"])
@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle"])
def test_overhang_trimmer(ctx, skip, which, algo):
    hf, hr = datasets.overhang_codes()
    kw = dict(hang_forward=hf if which != "reverse" else None, hang_reverse=hr if which != "forward" else None, hang_skip=skip)
    got, want, rep = run_both(ctx, pb.make_config(algo, **kw), datasets.overhang(2000))
    assert rep["ok"], rep


def test_overhang_strict_threshold_drops_pairs(ctx):
    hf, hr = datasets.overhang_codes()
    b = datasets.overhang(2000, seed=4)
    cfg = pb.make_config("simple_bayesian", hang_forward=hf, hang_reverse=hr, hang_threshold=np.log(0.9997), filters=[("short", 80), ("no_n", 0)])
    got, want, rep = run_both(ctx, cfg, b)
    assert rep["ok"], rep
    dropped = (got["results"]["status"] == 6).sum()
    assert 0 < dropped < b.n and got["counters"][pb.C_COUNT] == b.n - dropped


def test_overhang_mixed_lengths_and_long_reads(ctx):
    """the trimmer on the other read-length classes (the reverse read's suffix move depends on the class)"""
    hf, hr = datasets.overhang_codes()
    for b in (datasets.mixed(1500), datasets.long250(600), datasets.primers300(400)):
        got, want, rep = run_both(ctx, pb.make_config("simple_bayesian", hang_forward=hf[:9], hang_reverse=hr[:9]), b)
        assert rep["ok"], rep


# ---- the other host entry points: per-base log p as codes, caller-packed records -------------------------------------------
@pytest.mark.parametrize("algo,kw", [("simple_bayesian", {}), ("pear", {}), ("rdp_mle", dict(forward_trim=7, reverse_trim=3)),
                                     ("uparse", dict(minoverlap=10)), ("flash", {}), ("stitch", {})])
def test_codes_expand_to_the_same_doubles(ctx, algo, kw):
    """pb_assemble_host_codes + pb_posterior_table give bit for bit the per-base log p pb_assemble_host writes as doubles."""
    for b in (datasets.cfg1(3000), datasets.stress(1500), datasets.edge_cases(), datasets.mixed(1500)):
        cfg = pb.make_config(algo, **kw)
        a = ctx.assemble_host(cfg, b, want_nt=True, want_p=True)
        c = ctx.assemble_host_codes(cfg, b)
        assert np.array_equal(a["results"].view(np.uint8), c["results"].view(np.uint8))
        assert np.array_equal(a["seq_nt_packed"], c["seq_nt_packed"]) and np.array_equal(a["counters"], c["counters"])
        ok = a["results"]["status"] == 0
        mask = ok[:, None] & (np.arange(a["seq_stride"])[None, :] < a["results"]["seq_len"][:, None])
        assert np.array_equal(a["seq_p"][mask].view(np.uint64), c["seq_p"][mask].view(np.uint64))
        rep = compare(c, oracle_lib.assemble("port", cfg, b, seq_stride=c["seq_stride"]))
        assert rep["ok"], rep


def test_codes_refused_with_primers_after(ctx):
    fwd, rev = datasets.primer_codes()
    cfg = pb.make_config("simple_bayesian", forward_primer=fwd, reverse_primer=rev, post_primers=True)
    with pytest.raises(pb.PandaseqError):
        ctx.assemble_host_codes(cfg, datasets.primers300(50))


@pytest.mark.parametrize("algo", ["simple_bayesian", "pear", "rdp_mle"])
def test_packed_host_entry_point(ctx, algo):
    """Records packed on the host (pb_pack_host) through pb_assemble_host_packed: the same bytes as the AoS entry point, over
    several chunks of the host path (the chunk's offsets are rebased on the device)."""
    for b in (datasets.cfg1(5000), datasets.mixed(3000), datasets.edge_cases()):
        cfg = pb.make_config(algo)
        a = ctx.assemble_host(cfg, b, want_nt=True, want_p=False)
        reads, meta, max_len = pb.pack_host(b)
        c = ctx.assemble_host_packed(cfg, reads, meta, max_len, seq_stride=a["seq_stride"])
        assert np.array_equal(a["seq_nt_packed"], c["seq_nt_packed"]) and np.array_equal(a["counters"], c["counters"])
        # (the AoS entry point picks the kernel class by the longest read of each chunk, the packed one by the caller's
        # max_read_len: a chunk may run on another kernel, whose sums add the same terms in another order)
        for k in a["results"].dtype.names:
            if k in ("quality", "est_prob"):
                assert np.abs(a["results"][k] - c["results"][k]).max() <= 1e-9, k
            else:
                assert np.array_equal(a["results"][k], c["results"][k]), k
