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
