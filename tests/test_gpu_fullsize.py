"""BASELINE full size (10 M 2x150 pairs, config 2) on the GPU: properties that do not need the oracle to run 10 M pairs --
counter identities, run-to-run determinism, shard invariance -- plus exact parity on a random 20 k-pair sample."""
import numpy as np
import pytest

import oracle_lib
import pandaseq_b200 as pb
from parity import compare

pytestmark = pytest.mark.gpu
N = 10_000_000
CHUNK = 1_000_000


@pytest.fixture(scope="module")
def full(built):
    import torch
    from pandaseq_b200 import synth
    ctx = pb.Context(0)
    parts_r, parts_m, flats, base16, max_len = [], [], [], 0, 0
    for ci in range(N // CHUNK):
        rect = synth.generate_config(2, n=CHUNK, device="cuda", chunk_index=ci)
        f_data, f_off, r_data, r_off = rect.to_flat_tensors()
        reads, meta, ml, total = ctx.pack_device(f_data, f_off, r_data, r_off)
        meta[:, 0] += base16
        base16 += total // 16
        max_len = max(max_len, ml)
        parts_r.append(reads[:total])
        parts_m.append(meta)
        if ci in (0, 7):       # keep two chunks' raw reads on the host for the sampled oracle comparison
            flats.append((ci, synth.FlatBatch(f_data.cpu().numpy(), f_off.cpu().numpy().astype(np.uint64),
                                              r_data.cpu().numpy(), r_off.cpu().numpy().astype(np.uint64))))
    reads = torch.cat(parts_r + [torch.zeros(16, dtype=torch.uint8, device="cuda")])
    meta = torch.cat(parts_m)
    yield ctx, reads, meta, max_len, flats
    ctx.close()


def run(ctx, reads, meta, max_len, lo, hi):
    import torch
    n = hi - lo
    stride = (2 * max_len + 15) & ~15
    res = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    nt = torch.zeros((n, stride // 2), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.assemble_device(pb.make_config("simple_bayesian"), n, max_len, reads, meta[lo:hi].contiguous(), res, nt, None, stride, cnt)
    ctx.synchronize()
    return res, nt, cnt.cpu().numpy()


def digest(res, nt):
    """order-sensitive checksum of the result records and merged reads, computed on the device"""
    import torch
    r = res.view(torch.int64)
    w = torch.arange(1, r.shape[0] + 1, device=r.device, dtype=torch.int64)[:, None]
    a = int((r * w).sum().item())
    b = int((nt.view(torch.int64).sum(dim=1) * w[:, 0]).sum().item())
    return a, b


def test_full_size_properties(full):
    ctx, reads, meta, max_len, flats = full
    res, nt, cnt = run(ctx, reads, meta, max_len, 0, N)
    # counter identities (assembler.c:252-348: every pair bumps count and exactly one outcome counter)
    assert cnt[pb.C_COUNT] == N
    assert cnt[pb.C_OK] + cnt[pb.C_LOWQ] + cnt[pb.C_NOALGN] + cnt[pb.C_BADR] + cnt[pb.C_NOFP] + cnt[pb.C_NORP] == N
    assert cnt[pb.C_OVERLAPS:].sum() == cnt[pb.C_OK]
    assert cnt[pb.C_OK] > 0.999 * N
    r = res.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel()
    ok = r["status"] == 0
    assert int(ok.sum()) == cnt[pb.C_OK] and int(r["slow"].sum()) == cnt[pb.C_SLOW]
    assert r["overlap"][ok].max() == cnt[pb.C_LONGEST]
    assert np.array_equal(np.bincount(r["overlap"][ok], minlength=900)[:900], cnt[pb.C_OVERLAPS:pb.C_OVERLAPS + 900])
    assert ((r["seq_len"][ok].astype(int) + r["overlap"][ok]) == 300).all()          # F - o + R with no primers/trims
    # determinism
    res2, nt2, cnt2 = run(ctx, reads, meta, max_len, 0, N)
    assert digest(res, nt) == digest(res2, nt2) and np.array_equal(cnt, cnt2)
    # shard invariance: two halves == the whole (what --gpus 2 does, minus the second device)
    ra, na, ca = run(ctx, reads, meta, max_len, 0, N // 2)
    rb, nb, cb = run(ctx, reads, meta, max_len, N // 2, N)
    import torch
    assert digest(torch.cat([ra, rb]), torch.cat([na, nb])) == digest(res, nt)
    merged = ca + cb
    merged[pb.C_LONGEST] = max(ca[pb.C_LONGEST], cb[pb.C_LONGEST])
    assert np.array_equal(merged, cnt)
    # exact parity on a sample of the same pairs
    rng = np.random.default_rng(0)
    ntc = nt.cpu().numpy()
    for ci, flat in flats:
        idx = np.sort(rng.choice(CHUNK, 10_000, replace=False))
        sub = pb.synth.FlatBatch.from_pairs([(flat.pair(i)[0][:, 0], flat.pair(i)[0][:, 1], flat.pair(i)[1][:, 0], flat.pair(i)[1][:, 1]) for i in idx])
        want = oracle_lib.assemble("port", pb.make_config("simple_bayesian"), sub)
        g = ci * CHUNK + idx
        got = dict(results=r[g], seq_nt=pb.unpack_nt(ntc[g]), seq_p=None, counters=None)
        rep = compare(got, want, check_counters=False)
        assert rep["ok"], rep


def test_two_kernel_path_equals_general_kernel_at_full_size(full):
    """All 10 M pairs through the two-kernel path and through the general kernel alone: every integer field, the overlap score
    (bit pattern) and every byte of the merged-read rows identical; quality within 1e-12 (the two paths add the same per-base
    terms in different orders)."""
    import torch
    ctx, reads, meta, max_len, flats = full
    ctx.set_lanes(1)
    before = ctx.lanes_stats()
    res_a, nt_a, cnt_a = run(ctx, reads, meta, max_len, 0, N)
    after = ctx.lanes_stats()
    assert after[0] - before[0] == N and after[1] - before[1] < N // 1000
    ctx.set_lanes(0)
    res_b, nt_b, cnt_b = run(ctx, reads, meta, max_len, 0, N)
    assert ctx.lanes_stats()[0] == after[0]
    # and with the candidate overlaps from the hash join (pb::seed_kernel) instead of the diagonal sweep (pb_sweep.cuh):
    # `examined` depends on the exact candidate set.  (The sweep hands a few more pairs to the general kernel -- the ones
    # whose certificate fails -- so `quality` may differ in the last bits for those.)
    ctx.set_lanes(2)
    res_c, nt_c, cnt_c = run(ctx, reads, meta, max_len, 0, N)
    ctx.set_lanes(-1)
    a = res_a.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel()
    for res_x, nt_x, cnt_x in ((res_b, nt_b, cnt_b), (res_c, nt_c, cnt_c)):
        assert np.array_equal(cnt_a, cnt_x)
        assert torch.equal(nt_a, nt_x)
        b = res_x.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel()
        for k in ("status", "slow", "overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["est_prob"].view(np.uint64), b["est_prob"].view(np.uint64))
        assert np.abs(a["quality"] - b["quality"]).max() <= 1e-12


@pytest.mark.parametrize("algo,kw", [("simple_bayesian", {}), ("flash", {}), ("uparse", dict(minoverlap=12)),
                                     ("simple_bayesian", dict(maxoverlap=140, threshold=0.8, filters=[("short", 200), ("pear_test", (1.0, -1.0, 0.01))])),
                                     ("pear", {})])
def test_two_paths_agree_on_decorated_reads(built, algo, kw):
    """2 M pairs with N (0.1 % of bases), '#' tails (5 % of reads) and read-through inserts: about a third of the pairs are
    handed from the two-kernel path to the general kernel; the outcome must not depend on who assembled a pair."""
    import torch
    from pandaseq_b200 import synth
    ctx = pb.Context(0)
    n = 2_000_000
    rect = synth.generate(n, rl=(150, 150), tmpl=(120, 290), seed=4242, device="cuda", n_rate=0.001, btail_rate=0.05)
    f_data, f_off, r_data, r_off = rect.to_flat_tensors()
    reads, meta, ml, total = ctx.pack_device(f_data, f_off, r_data, r_off)
    reads = torch.cat([reads[:total], torch.zeros(16, dtype=torch.uint8, device="cuda")])
    cfg = pb.make_config(algo, **kw)
    stride = (2 * ml + 15) & ~15
    outs = []
    for mode in (1, 0, 2):
        ctx.set_lanes(mode)
        res = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
        nt = torch.zeros((n, stride // 2), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        before = ctx.lanes_stats()
        ctx.assemble_device(cfg, n, ml, reads, meta.contiguous(), res, nt, None, stride, cnt)
        ctx.synchronize()
        after = ctx.lanes_stats()
        outs.append((res.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel(), nt, cnt.cpu().numpy(), after[0] - before[0], after[1] - before[1]))
    ctx.close()
    (a, nt_a, cnt_a, lanes_a, deferred_a), (b, nt_b, cnt_b, lanes_b, _), (c, nt_c, cnt_c, lanes_c, _) = outs
    assert lanes_a == n and lanes_b == 0 and n // 10 < deferred_a < n and lanes_c == n
    # sweep seeding vs hash-join seeding
    assert torch.equal(nt_a, nt_c) and np.array_equal(cnt_a, cnt_c)
    for k in ("status", "slow", "overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset"):
        assert np.array_equal(a[k], c[k]), k
    assert np.abs(a["quality"] - c["quality"]).max() <= 1e-12 and np.abs(a["est_prob"] - c["est_prob"]).max() <= 1e-9
    assert np.array_equal(cnt_a, cnt_b)
    assert torch.equal(nt_a, nt_b)
    for k in ("status", "slow", "overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset"):
        assert np.array_equal(a[k], b[k]), k
    if algo == "pear":       # the lane kernel adds pear's terms in the reference's order, the general kernel with a shuffle tree
        assert np.abs(a["est_prob"] - b["est_prob"]).max() <= 1e-9
    else:
        assert np.array_equal(a["est_prob"].view(np.uint64), b["est_prob"].view(np.uint64))
    assert np.abs(a["quality"] - b["quality"]).max() <= 1e-12


# ---- BASELINE configs 3, 4 and 5 at 4 M pairs ---------------------------------------------------------------------------------------
N345 = 4_000_000


def _config_batch(ctx, cfg_id, n):
    """n pairs of a BASELINE config packed on the device, two of the 1 M-pair chunks kept on the host for the oracle sample"""
    import torch
    from pandaseq_b200 import synth
    parts_r, parts_m, flats, base16, max_len = [], [], [], 0, 0
    for ci in range(n // CHUNK):
        rect = synth.generate_config(cfg_id, n=CHUNK, device="cuda", chunk_index=ci)
        f_data, f_off, r_data, r_off = rect.to_flat_tensors()
        reads, meta, ml, total = ctx.pack_device(f_data, f_off, r_data, r_off)
        meta[:, 0] += base16
        base16 += total // 16
        max_len = max(max_len, ml)
        parts_r.append(reads[:total])
        parts_m.append(meta)
        if ci in (0, n // CHUNK - 1):
            flats.append((ci, synth.FlatBatch(f_data.cpu().numpy(), f_off.cpu().numpy().astype(np.uint64),
                                              r_data.cpu().numpy(), r_off.cpu().numpy().astype(np.uint64))))
    reads = torch.cat(parts_r + [torch.zeros(16, dtype=torch.uint8, device="cuda")])
    return reads, torch.cat(parts_m), max_len, flats


def _config_of(cfg_id):
    from pandaseq_b200 import synth
    c = synth.CONFIGS[cfg_id]
    kw = {}
    if c.get("primers"):
        kw = dict(forward_primer=synth.encode(synth.FWD_PRIMER),
                  reverse_primer=synth.encode("".join(synth._COMP[ch] for ch in synth.REV_PRIMER)))
    return pb.make_config(c["algo"], **kw)


@pytest.mark.parametrize("cfg_id", [3, 4, 5])
def test_configs_3_4_5_at_4m_pairs(built, cfg_id):
    """4 M pairs of BASELINE config 3 (2x250, pear), 4 (2x300, rdp_mle, both primers stripped) and 5 (2x(75-300) mixed lengths):
    counter identities, determinism, every available kernel path giving the same records (no status / overlap / base differs,
    i.e. no near-tie flips an arg-max or the threshold test between paths), and exact parity with the oracle on 10 k sampled pairs
    (status, overlap, offsets, mismatches, every merged base; quality and overlap score within 1e-6)."""
    import torch
    ctx = pb.Context(0)
    cfg = _config_of(cfg_id)
    reads, meta, max_len, flats = _config_batch(ctx, cfg_id, N345)
    stride = (2 * max_len + 15) & ~15

    def go():
        res = torch.zeros((N345, 32), dtype=torch.uint8, device="cuda")
        nt = torch.zeros((N345, stride // 2), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        before = ctx.lanes_stats()
        ctx.assemble_device(cfg, N345, max_len, reads, meta, res, nt, None, stride, cnt)
        ctx.synchronize()
        return res, nt, cnt.cpu().numpy(), ctx.lanes_stats()[0] - before[0]

    res, nt, cnt, on_lanes = go()
    assert cnt[pb.C_COUNT] == N345
    assert cnt[pb.C_OK] + cnt[pb.C_LOWQ] + cnt[pb.C_NOALGN] + cnt[pb.C_BADR] + cnt[pb.C_NOFP] + cnt[pb.C_NORP] == N345
    assert cnt[pb.C_OVERLAPS:].sum() == cnt[pb.C_OK] and cnt[pb.C_OK] > 0.9 * N345
    r = res.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel()
    ok = r["status"] == 0
    assert int(ok.sum()) == cnt[pb.C_OK] and int(r["slow"].sum()) == cnt[pb.C_SLOW]
    assert np.array_equal(np.bincount(r["overlap"][ok], minlength=900)[:900], cnt[pb.C_OVERLAPS:pb.C_OVERLAPS + 900])
    res2, nt2, cnt2, _ = go()
    assert digest(res, nt) == digest(res2, nt2) and np.array_equal(cnt, cnt2)
    # the other kernel paths on the same pairs: the general kernel alone, and (where the two-kernel path ran) the hash-join seeding
    for mode in ((0, 2) if on_lanes else ()):
        ctx.set_lanes(mode)
        res_x, nt_x, cnt_x, _ = go()
        ctx.set_lanes(-1)
        assert np.array_equal(cnt, cnt_x), mode
        assert torch.equal(nt, nt_x), mode
        x = res_x.cpu().numpy().view(pb.PAIR_RESULT_DTYPE).ravel()
        for k in ("status", "slow", "overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset"):
            assert np.array_equal(r[k], x[k]), (mode, k)
        assert np.abs(r["quality"] - x["quality"]).max() <= 1e-9 and np.abs(r["est_prob"][ok] - x["est_prob"][ok]).max() <= 1e-9, mode
    rng = np.random.default_rng(cfg_id)
    ntc = nt.cpu().numpy()
    for ci, flat in flats:
        idx = np.sort(rng.choice(CHUNK, 5_000, replace=False))
        sub = pb.synth.FlatBatch.from_pairs([(flat.pair(i)[0][:, 0], flat.pair(i)[0][:, 1], flat.pair(i)[1][:, 0], flat.pair(i)[1][:, 1]) for i in idx])
        want = oracle_lib.assemble("port", cfg, sub)
        g = ci * CHUNK + idx
        got = dict(results=r[g], seq_nt=pb.unpack_nt(ntc[g]), seq_p=None, counters=None)
        rep = compare(got, want, check_counters=False)
        assert rep["ok"], (cfg_id, rep)
    ctx.close()
