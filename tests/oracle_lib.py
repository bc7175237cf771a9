"""ctypes access to the CHECKERS: oracle/libpanda_oracle.so (our CPU restatement) and, when it has
been built, oracle/_ref/libref_harness.so (the unmodified reference).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from pandaseq_b200 import PB_MAX_LEN, PB_NCOUNTERS, PbConfig, PbTables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_PATH = os.path.join(ROOT, "oracle", "libpanda_oracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")


class FlatOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("status", "slow", "overlap", "seq_len", "mismatches", "degenerates", "examined",
                                          "fwd_offset", "rev_offset", "quality", "est_prob", "seq_nt", "seq_p")] + \
               [("seq_stride", C.c_int64), ("counters", C.c_void_p)]


_FIELDS = dict(status=np.uint8, slow=np.uint8, overlap=np.int32, seq_len=np.int32, mismatches=np.int32, degenerates=np.int32,
               examined=np.int32, fwd_offset=np.int32, rev_offset=np.int32, quality=np.float64, est_prob=np.float64)


def _load(path, prefix):
    L = C.CDLL(path)
    fn = getattr(L, prefix + "_assemble_flat")
    fn.argtypes = [C.POINTER(PbConfig), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FlatOut), C.c_int]
    fn.restype = C.c_int
    off = getattr(L, prefix + "_compute_offset_qual")
    off.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
    off.restype = C.c_size_t
    return L, fn, off


_cache = {}


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def _get(which):
    if which not in _cache:
        _cache[which] = _load(PORT_PATH, "po") if which == "port" else _load(REF_PATH, "ref")
    return _cache[which]


def assemble(which: str, cfg: PbConfig, batch, *, want_seq=True, threads=1, seq_stride=None):
    """which: 'port' (oracle/panda_oracle.c) or 'ref' (the compiled reference).  Returns dict of numpy arrays."""
    _, fn, _ = _get(which)
    n = batch.n
    if seq_stride is None:
        fl, rl = batch.lengths()
        seq_stride = int((fl + rl).max()) if n else 0
        seq_stride = (seq_stride + 15) & ~15
    out = {k: np.zeros(n, dtype=dt) for k, dt in _FIELDS.items()}
    out["seq_nt"] = np.zeros((n, seq_stride), np.uint8) if want_seq else None
    out["seq_p"] = np.zeros((n, seq_stride), np.float64) if want_seq else None
    out["counters"] = np.zeros(PB_NCOUNTERS, np.int64)
    fo = FlatOut()
    for k in list(_FIELDS) + ["seq_nt", "seq_p", "counters"]:
        setattr(fo, k, None if out[k] is None else out[k].ctypes.data)
    fo.seq_stride = seq_stride
    f_data, r_data = np.ascontiguousarray(batch.f_data), np.ascontiguousarray(batch.r_data)
    f_off, r_off = np.ascontiguousarray(batch.f_off, dtype=np.uint64), np.ascontiguousarray(batch.r_off, dtype=np.uint64)
    rc = fn(C.byref(cfg), n, f_data.ctypes.data, f_off.ctypes.data, r_data.ctypes.data, r_off.ctypes.data, C.byref(fo), int(threads))
    if rc != 0:
        raise RuntimeError(f"{which}_assemble_flat returned {rc}")
    out["seq_stride"] = seq_stride
    return out


def compute_offset(which: str, threshold_log: float, penalty: float, reverse: bool, read: np.ndarray, needle: bytes) -> int:
    _, _, off = _get(which)
    read = np.ascontiguousarray(read, dtype=np.uint8)
    return int(off(threshold_log, penalty, int(reverse), read.ctypes.data, len(read), needle, len(needle)))


def tables(which: str) -> dict:
    L, _, _ = _get(which)
    if which == "port":
        L.po_get_tables.restype = C.POINTER(PbTables)
        return L.po_get_tables().contents.as_dict()
    t = PbTables()
    L.ref_get_tables.argtypes = [C.POINTER(PbTables)]
    L.ref_get_tables(C.byref(t))
    return t.as_dict()


# ---- the stages either side of the hot path: FASTQ text in, FASTA/FASTQ text out ---------------------------
SEQID_DTYPE = np.dtype([("instrument", "S100"), ("run", "S100"), ("flowcell", "S100"), ("lane", "<i4"), ("tile", "<i4"),
                        ("x", "<i4"), ("y", "<i4"), ("tag", "S50"), ("_pad", "V2")])
assert SEQID_DTYPE.itemsize == 368
FQ_ERRORS = {0: "OK", 1: "BADID", 2: "NOTPAIRED", 3: "EOF", 4: "BADNT", 5: "READLEN", 6: "BADSEQ", 7: "NOQUAL"}
TAG_PRESENT, TAG_ABSENT, TAG_OPTIONAL = 0, 1, 2


class FastqOut(C.Structure):
    _fields_ = [("n", C.c_size_t), ("records", C.c_size_t), ("error", C.c_int), ("ids", C.c_void_p), ("f_data", C.c_void_p),
                ("f_off", C.c_void_p), ("r_data", C.c_void_p), ("r_off", C.c_void_p)]


def _io(which):
    L = _get(which)[0]
    pre = "po" if which == "port" else "ref"
    return L, pre


def fastq_parse(which: str, fwd: bytes, rev: bytes, *, qualmin=33, policy=TAG_PRESENT, max_pairs=None, max_read=0):
    """-> dict(n, error, ids (SEQID_DTYPE), batch (FlatBatch), records)"""
    from pandaseq_b200.synth import FlatBatch
    L, pre = _io(which)
    fn = getattr(L, pre + "_fastq_parse")
    fn.restype = C.c_int
    if max_pairs is None:
        max_pairs = fwd.count(b"\n") // 4 + 1
    cap = max_pairs * PB_MAX_LEN
    ids = np.zeros(max_pairs, dtype=SEQID_DTYPE)
    f_data, r_data = np.zeros((cap, 2), np.uint8), np.zeros((cap, 2), np.uint8)
    f_off, r_off = np.zeros(max_pairs + 1, np.uint64), np.zeros(max_pairs + 1, np.uint64)
    out = FastqOut(0, 0, 0, ids.ctypes.data, f_data.ctypes.data, f_off.ctypes.data, r_data.ctypes.data, r_off.ctypes.data)
    args = [C.c_char_p(fwd), C.c_size_t(len(fwd)), C.c_char_p(rev), C.c_size_t(len(rev)), C.c_int(qualmin), C.c_int(policy),
            C.c_size_t(max_pairs), C.byref(out)]
    if which != "port":
        args.append(C.c_size_t(max_read))
    fn(*args)
    n = int(out.n)
    batch = FlatBatch(f_data[:int(f_off[n])].copy(), f_off[:n + 1].copy(), r_data[:int(r_off[n])].copy(), r_off[:n + 1].copy())
    return dict(n=n, error=int(out.error), ids=ids[:n].copy(), batch=batch, records=int(out.records))


def seqid_parse(which: str, text: bytes, policy=TAG_PRESENT):
    L, pre = _io(which)
    fn = getattr(L, pre + "_seqid_parse")
    fn.restype = C.c_int
    ident = np.zeros(1, dtype=SEQID_DTYPE)
    fmt = C.c_int(0)
    rc = fn(C.c_void_p(ident.ctypes.data), C.c_char_p(text), C.c_int(policy), C.byref(fmt))
    return int(rc), int(fmt.value), ident[0]


def result_phred(which: str, p: float) -> int:
    L, pre = _io(which)
    fn = getattr(L, pre + "_result_phred")
    fn.restype = C.c_char
    fn.argtypes = [C.c_double]
    return fn(float(p))[0]


def format_flat(which: str, fastq: bool, ids, status, quality, seq_len, seq_nt, seq_p, seq_stride) -> bytes:
    L, pre = _io(which)
    fn = getattr(L, pre + "_format_flat")
    fn.restype = C.c_size_t
    n = len(status)
    ids = np.ascontiguousarray(ids)
    status = np.ascontiguousarray(status, dtype=np.uint8)
    quality = np.ascontiguousarray(quality, dtype=np.float64)
    seq_len = np.ascontiguousarray(seq_len, dtype=np.int32)
    seq_nt = np.ascontiguousarray(seq_nt, dtype=np.uint8)
    seq_p = None if seq_p is None else np.ascontiguousarray(seq_p, dtype=np.float64)
    dst = np.zeros(n * (2 * (2 * PB_MAX_LEN) + 512) + 16, dtype=np.uint8)
    total = fn(C.c_void_p(dst.ctypes.data), C.c_int(int(fastq)), C.c_size_t(n), C.c_void_p(ids.ctypes.data), C.c_void_p(status.ctypes.data),
               C.c_void_p(quality.ctypes.data), C.c_void_p(seq_len.ctypes.data), C.c_void_p(seq_nt.ctypes.data),
               C.c_void_p(seq_p.ctypes.data) if seq_p is not None else None, C.c_int64(seq_stride))
    return dst[:total].tobytes()
