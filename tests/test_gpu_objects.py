"""The reference-shaped object API (panda_assembler_* / panda_algorithm_*) on the GPU, called through ctypes, and a plain C
program compiled against include/pandaseq_b200.h: per-pair assemble, pull-source next(), batch call, counters, callbacks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import datasets
import oracle_lib
import pandaseq_b200 as pb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Qual(C.Structure):
    _fields_ = [("nt", C.c_char), ("qual", C.c_char)]


class Result(C.Structure):
    _fields_ = [("nt", C.c_char), ("p", C.c_double)]


class SeqId(C.Structure):
    _fields_ = [("instrument", C.c_char * 100), ("run", C.c_char * 100), ("flowcell", C.c_char * 100),
                ("lane", C.c_int), ("tile", C.c_int), ("x", C.c_int), ("y", C.c_int), ("tag", C.c_char * 50)]


class ResultSeq(C.Structure):
    _fields_ = [("quality", C.c_double), ("degenerates", C.c_size_t), ("name", SeqId), ("sequence", C.POINTER(Result)),
                ("sequence_length", C.c_size_t), ("forward", C.c_void_p), ("forward_length", C.c_size_t),
                ("reverse", C.c_void_p), ("reverse_length", C.c_size_t), ("forward_offset", C.c_size_t),
                ("reverse_offset", C.c_size_t), ("overlap_mismatches", C.c_size_t), ("overlaps_examined", C.c_size_t),
                ("overlap", C.c_size_t), ("estimated_overlap_probability", C.c_double)]


NEXT = C.CFUNCTYPE(C.c_bool, C.POINTER(SeqId), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p)
OUTPUT = C.CFUNCTYPE(C.c_bool, C.POINTER(ResultSeq), C.c_void_p)
FAIL = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(SeqId), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p)


@pytest.fixture(scope="module")
def L(built):
    lib = pb.lib()
    assert C.sizeof(SeqId) == 368 and C.sizeof(Result) == 16 and C.sizeof(Qual) == 2
    vp = C.c_void_p
    lib.panda_assembler_new.restype = vp
    lib.panda_assembler_new.argtypes = [vp, vp, vp, vp]
    lib.panda_assembler_assemble.restype = C.POINTER(ResultSeq)
    lib.panda_assembler_assemble.argtypes = [vp, C.POINTER(SeqId), vp, C.c_size_t, vp, C.c_size_t]
    lib.panda_assembler_next.restype = C.POINTER(ResultSeq)
    lib.panda_assembler_next.argtypes = [vp]
    lib.panda_assembler_assemble_batch.restype = C.c_size_t
    lib.panda_assembler_assemble_batch.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp]
    for name in ("count", "ok_count", "low_quality_count", "failed_alignment_count", "bad_read_count", "slow_count",
                 "no_forward_primer_count", "no_reverse_primer_count"):
        fn = getattr(lib, "panda_assembler_get_" + name)
        fn.restype, fn.argtypes = C.c_long, [vp]
    lib.panda_assembler_get_overlap_count.restype, lib.panda_assembler_get_overlap_count.argtypes = C.c_long, [vp, C.c_size_t]
    lib.panda_assembler_get_longest_overlap.restype, lib.panda_assembler_get_longest_overlap.argtypes = C.c_size_t, [vp]
    lib.panda_assembler_get_threshold.restype, lib.panda_assembler_get_threshold.argtypes = C.c_double, [vp]
    lib.panda_assembler_set_threshold.argtypes = [vp, C.c_double]
    lib.panda_assembler_set_minimum_overlap.argtypes = [vp, C.c_int]
    lib.panda_assembler_set_maximum_overlap.argtypes = [vp, C.c_int]
    lib.panda_assembler_get_minimum_overlap.argtypes = [vp]
    lib.panda_assembler_set_algorithm.argtypes = [vp, vp]
    lib.panda_assembler_set_forward_primer.argtypes = [vp, vp, C.c_size_t]
    lib.panda_assembler_set_reverse_primer.argtypes = [vp, vp, C.c_size_t]
    lib.panda_assembler_set_forward_trim.argtypes = [vp, C.c_size_t]
    lib.panda_assembler_get_forward_trim.restype, lib.panda_assembler_get_forward_trim.argtypes = C.c_size_t, [vp]
    lib.panda_assembler_get_forward_primer.restype, lib.panda_assembler_get_forward_primer.argtypes = vp, [vp, C.POINTER(C.c_size_t)]
    lib.panda_assembler_copy_configuration.argtypes = [vp, vp]
    lib.panda_assembler_set_fail_alignment.argtypes = [vp, vp, vp, vp]
    lib.panda_assembler_unref.argtypes = [vp]
    lib.panda_algorithm_unref.argtypes = [vp]
    for name in ("panda_algorithm_pear_new", "panda_algorithm_rdp_mle_new", "panda_algorithm_simple_bayes_new", "panda_algorithm_flash_new"):
        getattr(lib, name).restype = vp
    lib.panda_compute_offset_qual.restype = C.c_size_t
    lib.panda_compute_offset_qual.argtypes = [C.c_double, C.c_double, C.c_bool, vp, C.c_size_t, vp, C.c_size_t]
    return lib


def check_result(res, want, i):
    r = res.contents
    assert r.overlap == want["overlap"][i] and r.sequence_length == want["seq_len"][i]
    assert r.overlap_mismatches == want["mismatches"][i] and r.degenerates == want["degenerates"][i]
    assert r.overlaps_examined == want["examined"][i]
    assert r.forward_offset == want["fwd_offset"][i] and r.reverse_offset == want["rev_offset"][i]
    assert abs(r.quality - want["quality"][i]) <= 1e-6
    assert abs(r.estimated_overlap_probability - want["est_prob"][i]) <= 1e-6
    n = r.sequence_length
    nt = np.array([r.sequence[k].nt[0] for k in range(n)], dtype=np.uint8)
    p = np.array([r.sequence[k].p for k in range(n)])
    assert np.array_equal(nt, want["seq_nt"][i, :n])
    assert np.abs(p - want["seq_p"][i, :n]).max() <= 1e-6


def test_assemble_one_pair_at_a_time(L):
    b = datasets.cfg1(120)
    want = oracle_lib.assemble("port", pb.make_config("pear"), b)
    a = L.panda_assembler_new(None, None, None, None)
    assert a, pb.lib().pb_last_error()
    algo = L.panda_algorithm_pear_new()
    L.panda_assembler_set_algorithm(a, algo)
    L.panda_algorithm_unref(algo)
    sid = SeqId()
    for i in range(b.n):
        f, r = b.pair(i)
        f, r = np.ascontiguousarray(f), np.ascontiguousarray(r)
        res = L.panda_assembler_assemble(a, C.byref(sid), f.ctypes.data, len(f), r.ctypes.data, len(r))
        assert bool(res) == (want["status"][i] == 0)
        if res:
            check_result(res, want, i)
            assert res.contents.forward == f.ctypes.data and res.contents.forward_length == len(f)
    assert L.panda_assembler_get_count(a) == b.n
    assert L.panda_assembler_get_ok_count(a) == want["counters"][pb.C_OK]
    assert L.panda_assembler_get_low_quality_count(a) == want["counters"][pb.C_LOWQ]
    assert L.panda_assembler_get_failed_alignment_count(a) == want["counters"][pb.C_NOALGN]
    assert L.panda_assembler_get_slow_count(a) == want["counters"][pb.C_SLOW]
    assert L.panda_assembler_get_longest_overlap(a) == want["counters"][pb.C_LONGEST]
    for ov in range(0, 900, 37):
        assert L.panda_assembler_get_overlap_count(a, ov) == want["counters"][pb.C_OVERLAPS + ov]
    assert L.panda_assembler_get_overlap_count(a, 900) == -1
    L.panda_assembler_unref(a)


def test_pull_source_next_and_callbacks(L, monkeypatch):
    monkeypatch.setenv("PANDASEQ_B200_NEXT_BATCH", "256")      # 700 pairs -> three device batches behind next()
    b = datasets.stress(700)
    want = oracle_lib.assemble("port", pb.make_config("simple_bayesian", minoverlap=10, threshold=0.7), b)
    state = {"i": 0, "keep": []}

    def nxt(idp, fp, flp, rp, rlp, _):
        i = state["i"]
        if i >= b.n:
            return False
        f, r = b.pair(i)
        f, r = np.ascontiguousarray(f), np.ascontiguousarray(r)
        state["keep"] = [f, r]                      # valid only until the next call, as the reference's sources are
        idp.contents.x = i
        fp[0], flp[0], rp[0], rlp[0] = f.ctypes.data, len(f), r.ctypes.data, len(r)
        state["i"] = i + 1
        return True

    failed = []
    cb_next, cb_fail = NEXT(nxt), FAIL(lambda a, idp, f, fl, r, rl, u: failed.append(idp.contents.x))
    a = L.panda_assembler_new(C.cast(cb_next, C.c_void_p), None, None, None)
    L.panda_assembler_set_minimum_overlap(a, 10)
    L.panda_assembler_set_threshold(a, 0.7)
    L.panda_assembler_set_fail_alignment(a, C.cast(cb_fail, C.c_void_p), None, None)
    assert abs(L.panda_assembler_get_threshold(a) - 0.7) < 1e-15 and L.panda_assembler_get_minimum_overlap(a) == 10
    got = []
    while True:
        res = L.panda_assembler_next(a)
        if not res:
            break
        i = res.contents.name.x
        got.append(i)
        check_result(res, want, i)
    assert got == np.nonzero(want["status"] == 0)[0].tolist()        # input order
    assert failed == np.nonzero(want["status"] == 4)[0].tolist()
    assert L.panda_assembler_get_count(a) == b.n
    L.panda_assembler_unref(a)


def test_batch_call_with_primers_and_copy_configuration(L):
    fwd, rev = datasets.primer_codes()
    b = datasets.primers300(300)
    want = oracle_lib.assemble("port", pb.make_config("rdp_mle", forward_primer=fwd, reverse_primer=rev), b)
    proto = L.panda_assembler_new(None, None, None, None)
    algo = L.panda_algorithm_rdp_mle_new()
    L.panda_assembler_set_algorithm(proto, algo)
    L.panda_algorithm_unref(algo)
    L.panda_assembler_set_forward_trim(proto, 5)
    L.panda_assembler_set_forward_primer(proto, fwd.ctypes.data, len(fwd))       # a primer clears the trim (assembler_support.c:201-213)
    L.panda_assembler_set_reverse_primer(proto, rev.ctypes.data, len(rev))
    assert L.panda_assembler_get_forward_trim(proto) == 0
    a = L.panda_assembler_new(None, None, None, None)
    L.panda_assembler_copy_configuration(a, proto)
    ln = C.c_size_t()
    assert L.panda_assembler_get_forward_primer(a, C.byref(ln)) and ln.value == len(fwd)
    fs = [np.ascontiguousarray(b.pair(i)[0]) for i in range(b.n)]
    rs = [np.ascontiguousarray(b.pair(i)[1]) for i in range(b.n)]
    fp = (C.c_void_p * b.n)(*[x.ctypes.data for x in fs])
    rp = (C.c_void_p * b.n)(*[x.ctypes.data for x in rs])
    fl = (C.c_size_t * b.n)(*[len(x) for x in fs])
    rl = (C.c_size_t * b.n)(*[len(x) for x in rs])
    ids = (SeqId * b.n)()
    for i in range(b.n):
        ids[i].x = i
    seen = []

    def out(res, _):
        i = res.contents.name.x
        seen.append(i)
        check_result(res, want, i)
        return True
    cb = OUTPUT(out)
    n_ok = L.panda_assembler_assemble_batch(a, b.n, ids, fp, fl, rp, rl, C.cast(cb, C.c_void_p), None)
    assert n_ok == int((want["status"] == 0).sum()) and seen == np.nonzero(want["status"] == 0)[0].tolist()
    assert L.panda_assembler_get_no_forward_primer_count(a) == want["counters"][pb.C_NOFP]
    L.panda_assembler_unref(a)
    L.panda_assembler_unref(proto)


def test_compute_offset_qual_entry_point(L):
    fwd, _ = datasets.primer_codes()
    b = datasets.primers300(40)
    thr = float(np.log(0.6))
    for i in range(0, 40, 3):
        f, r = b.pair(i)
        for read in (np.ascontiguousarray(f), np.ascontiguousarray(r)):
            for rv in (False, True):
                want = oracle_lib.compute_offset("port", thr, 0.0, rv, read, bytes(fwd.tolist()))
                got = L.panda_compute_offset_qual(thr, 0.0, rv, read.ctypes.data, len(read), fwd.ctypes.data, len(fwd))
                assert got == want


def test_c_program_against_the_header(L, tmp_path):
    """tests/c/dropin_demo.c: compiled with gcc against include/pandaseq_b200.h, linked with the library, run on the GPU."""
    exe = str(tmp_path / "dropin_demo")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "dropin_demo.c"),
                    "-L", os.path.join(ROOT, "pandaseq_b200"), "-lpandaseq_b200", "-Wl,-rpath," + os.path.join(ROOT, "pandaseq_b200"),
                    "-lpthread", "-o", exe], check=True)
    b = datasets.cfg1(2500)
    path = str(tmp_path / "pairs.bin")
    with open(path, "wb") as f:
        f.write(np.array([b.n, len(b.f_data), len(b.r_data)], dtype=np.uint64).tobytes())
        f.write(b.f_off.tobytes()); f.write(b.r_off.tobytes()); f.write(b.f_data.tobytes()); f.write(b.r_data.tobytes())
    for algo in ("simple_bayesian", "rdp_mle"):
        want = oracle_lib.assemble("port", pb.make_config(algo), b)
        run = subprocess.run([exe, path, algo], capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        stat = dict(line.split("\t")[1:3] for line in run.stderr.strip().split("\n") if line.startswith("STAT"))
        assert int(stat["READS"]) == b.n and int(stat["OK"]) == want["counters"][pb.C_OK]
        assert int(stat["NOALGN"]) == want["counters"][pb.C_NOALGN] and int(stat["LOWQ"]) == want["counters"][pb.C_LOWQ]
        lines = run.stdout.strip().split("\n")
        ok = np.nonzero(want["status"] == 0)[0]
        assert len(lines) == 2 * len(ok)
        letters = np.frombuffer(b"NACMGRSVTWYHKDBN", dtype=np.uint8)
        for k, i in enumerate(ok[:200]):
            assert lines[2 * k].startswith(f">pair{i};overlap={want['overlap'][i]};")
            assert lines[2 * k + 1] == letters[want["seq_nt"][i, :want["seq_len"][i]]].tobytes().decode()


def test_c_program_with_a_pool_of_workers(L, tmp_path):
    """panda_run_pool(threads = 5): five workers, each on GPU (worker mod visible GPUs), pull batches of 300 pairs from the one
    source and hand their results out concurrently.  The output order is unspecified (pandaseq.1:208); the set of records and
    the merged STAT counters are the single-threaded run's."""
    exe = str(tmp_path / "dropin_demo")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "dropin_demo.c"),
                    "-L", os.path.join(ROOT, "pandaseq_b200"), "-lpandaseq_b200", "-Wl,-rpath," + os.path.join(ROOT, "pandaseq_b200"),
                    "-lpthread", "-o", exe], check=True)
    b = datasets.cfg1(4000)
    path = str(tmp_path / "pairs.bin")
    with open(path, "wb") as f:
        f.write(np.array([b.n, len(b.f_data), len(b.r_data)], dtype=np.uint64).tobytes())
        f.write(b.f_off.tobytes()); f.write(b.r_off.tobytes()); f.write(b.f_data.tobytes()); f.write(b.r_data.tobytes())
    env = dict(os.environ, PANDASEQ_B200_NEXT_BATCH="300")
    runs = {}
    for threads in (1, 5):
        run = subprocess.run([exe, path, "simple_bayesian", str(threads)], capture_output=True, text=True, env=env)
        assert run.returncode == 0, run.stderr
        lines = run.stdout.strip().split("\n")
        recs = sorted(zip(lines[0::2], lines[1::2]), key=lambda t: int(t[0][5:t[0].index(";")]))
        runs[threads] = (recs, [ln for ln in run.stderr.strip().split("\n") if ln.startswith("STAT")])
    assert runs[1][0] == runs[5][0] and len(runs[1][0]) > 3900
    assert runs[1][1] == runs[5][1]


def test_a_call_between_two_next_calls_leaves_the_stream_alone(L, monkeypatch):
    """assembler.c:350-383: panda_assembler_assemble() and panda_assembler_next() are independent.  next() works through a
    prefetched batch here; a single-pair call in between uses its own staging and must not drop what is left of that batch."""
    monkeypatch.setenv("PANDASEQ_B200_NEXT_BATCH", "128")
    b = datasets.cfg1(300)
    want = oracle_lib.assemble("port", pb.make_config("simple_bayesian"), b)
    state = {"i": 0, "keep": []}

    def nxt(idp, fp, flp, rp, rlp, _):
        i = state["i"]
        if i >= b.n:
            return False
        f, r = b.pair(i)
        f, r = np.ascontiguousarray(f), np.ascontiguousarray(r)
        state["keep"] = [f, r]
        idp.contents.x = i
        fp[0], flp[0], rp[0], rlp[0] = f.ctypes.data, len(f), r.ctypes.data, len(r)
        state["i"] = i + 1
        return True

    cb = NEXT(nxt)
    a = L.panda_assembler_new(C.cast(cb, C.c_void_p), None, None, None)
    got, sid, extra = [], SeqId(), 0
    ef, er = (np.ascontiguousarray(x) for x in b.pair(7))
    while True:
        res = L.panda_assembler_next(a)
        if not res:
            break
        i = res.contents.name.x
        got.append(i)
        check_result(res, want, i)
        if len(got) % 10 == 3:          # in the middle of a batch
            one = L.panda_assembler_assemble(a, C.byref(sid), ef.ctypes.data, len(ef), er.ctypes.data, len(er))
            assert bool(one) == (want["status"][7] == 0)
            if one:
                check_result(one, want, 7)
            extra += 1
    assert got == np.nonzero(want["status"] == 0)[0].tolist()
    assert L.panda_assembler_get_count(a) == b.n + extra
    L.panda_assembler_unref(a)


# ---- modules with host callbacks (pandaseq-module.h, module.c:124-154) -------------------------------------------------
PRECHECK = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.POINTER(SeqId), C.POINTER(Qual), C.c_size_t, C.POINTER(Qual), C.c_size_t, C.c_void_p)
CHECK = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.POINTER(ResultSeq), C.c_void_p)
MODCB = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def _module_api(L):
    vp = C.c_void_p
    L.panda_module_new.restype, L.panda_module_new.argtypes = vp, [C.c_char_p, vp, vp, vp, vp]
    L.panda_module_unref.argtypes = [vp]
    L.panda_module_get_name.restype, L.panda_module_get_name.argtypes = C.c_char_p, [vp]
    L.panda_assembler_add_module.restype, L.panda_assembler_add_module.argtypes = C.c_bool, [vp, vp]
    L.panda_assembler_foreach_module.restype, L.panda_assembler_foreach_module.argtypes = C.c_bool, [vp, vp, vp]


def _expected_with_modules(b, want, pre_reject, check_reject):
    """What assemble_seq does with a pre-check and a check module (assembler.c:252-348): the pre-check sees every pair with two
    bases per read, the check every pair that passed the quality threshold; only pairs passing both are counted as OK."""
    fl, rl = b.lengths()
    seen_pre = (fl >= 2) & (rl >= 2)
    pre = np.array([seen_pre[i] and pre_reject(i) for i in range(b.n)])
    okdev = (want["status"] == 0) & ~pre
    chk = np.array([okdev[i] and check_reject(i) for i in range(b.n)])
    return pre, chk, okdev & ~chk


def test_modules_precheck_and_check_behind_next_batch_and_single(L, monkeypatch):
    _module_api(L)
    monkeypatch.setenv("PANDASEQ_B200_NEXT_BATCH", "300")
    b = datasets.cfg1(1000)
    want = oracle_lib.assemble("port", pb.make_config("simple_bayesian"), b)
    pre_reject = lambda i: int(b.pair(i)[0][0, 0]) == 1                  # forward read starts with A
    check_reject = lambda i: int(want["seq_len"][i]) % 3 == 0
    pre, chk, ok = _expected_with_modules(b, want, pre_reject, check_reject)
    assert pre.sum() > 50 and chk.sum() > 50 and ok.sum() > 300
    calls = {"pre": 0, "chk": 0}

    def precheck(logger, idp, f, fl, r, rl, user):
        calls["pre"] += 1
        return f[0].nt[0] != 1

    def check(logger, res, user):
        calls["chk"] += 1
        return res.contents.sequence_length % 3 != 0

    cb_pre, cb_chk = PRECHECK(precheck), CHECK(check)
    m_pre = L.panda_module_new(b"starts_with_a", None, C.cast(cb_pre, C.c_void_p), None, None)
    m_chk = L.panda_module_new(b"multiple_of_three", C.cast(cb_chk, C.c_void_p), None, None, None)
    assert m_pre and m_chk and L.panda_module_get_name(m_chk) == b"multiple_of_three"
    assert not L.panda_module_new(b"nothing", None, None, None, None)

    def module_counts(a):
        out = []
        cb = MODCB(lambda asm, mod, rejected, data: out.append((L.panda_module_get_name(mod), rejected)) or True)
        assert L.panda_assembler_foreach_module(a, C.cast(cb, C.c_void_p), None)
        return out

    def check_counters(a):
        assert L.panda_assembler_get_count(a) == b.n
        assert L.panda_assembler_get_ok_count(a) == int(ok.sum())
        assert L.panda_assembler_get_low_quality_count(a) == int(((want["status"] == 5) & ~pre).sum())
        assert L.panda_assembler_get_failed_alignment_count(a) == int(((want["status"] == 4) & ~pre).sum())
        assert L.panda_assembler_get_longest_overlap(a) == int(want["overlap"][ok].max())
        hist = np.bincount(want["overlap"][ok], minlength=900)
        for ov in range(0, 900, 7):
            assert L.panda_assembler_get_overlap_count(a, ov) == hist[ov]
        assert module_counts(a) == [(b"starts_with_a", int(pre.sum())), (b"multiple_of_three", int(chk.sum()))]

    # 1. pull source + next(), modules copied from a prototype assembler (assembler_support.c:123-125)
    proto = L.panda_assembler_new(None, None, None, None)
    assert L.panda_assembler_add_module(proto, m_pre) and L.panda_assembler_add_module(proto, m_chk)
    state = {"i": 0, "keep": []}

    def nxt(idp, fp, flp, rp, rlp, _):
        i = state["i"]
        if i >= b.n:
            return False
        f, r = b.pair(i)
        f, r = np.ascontiguousarray(f), np.ascontiguousarray(r)
        state["keep"] = [f, r]
        idp.contents.x = i
        fp[0], flp[0], rp[0], rlp[0] = f.ctypes.data, len(f), r.ctypes.data, len(r)
        state["i"] = i + 1
        return True

    cb_next = NEXT(nxt)
    a = L.panda_assembler_new(C.cast(cb_next, C.c_void_p), None, None, None)
    L.panda_assembler_copy_configuration(a, proto)
    got = []
    while True:
        res = L.panda_assembler_next(a)
        if not res:
            break
        got.append(res.contents.name.x)
        check_result(res, want, got[-1])
    assert got == np.nonzero(ok)[0].tolist()
    check_counters(a)
    fl, rl = b.lengths()
    assert calls["pre"] == int(((fl >= 2) & (rl >= 2)).sum()) and calls["chk"] == int(((want["status"] == 0) & ~pre).sum())
    L.panda_assembler_unref(a)

    # 2. the batch call
    a = L.panda_assembler_new(None, None, None, None)
    L.panda_assembler_copy_configuration(a, proto)
    fs = [np.ascontiguousarray(b.pair(i)[0]) for i in range(b.n)]
    rs = [np.ascontiguousarray(b.pair(i)[1]) for i in range(b.n)]
    fp = (C.c_void_p * b.n)(*[x.ctypes.data for x in fs])
    rp = (C.c_void_p * b.n)(*[x.ctypes.data for x in rs])
    fla = (C.c_size_t * b.n)(*[len(x) for x in fs])
    rla = (C.c_size_t * b.n)(*[len(x) for x in rs])
    ids = (SeqId * b.n)()
    for i in range(b.n):
        ids[i].x = i
    seen = []
    cb_out = OUTPUT(lambda res, _: seen.append(res.contents.name.x) or True)
    assert L.panda_assembler_assemble_batch(a, b.n, ids, fp, fla, rp, rla, C.cast(cb_out, C.c_void_p), None) == int(ok.sum())
    assert seen == np.nonzero(ok)[0].tolist()
    check_counters(a)
    L.panda_assembler_unref(a)

    # 3. one pair at a time
    a = L.panda_assembler_new(None, None, None, None)
    L.panda_assembler_copy_configuration(a, proto)
    sid = SeqId()
    for i in range(0, b.n, 5):
        res = L.panda_assembler_assemble(a, C.byref(sid), fs[i].ctypes.data, len(fs[i]), rs[i].ctypes.data, len(rs[i]))
        assert bool(res) == bool(ok[i])
    L.panda_assembler_unref(a)
    L.panda_assembler_unref(proto)
    L.panda_module_unref(m_pre)
    L.panda_module_unref(m_chk)
