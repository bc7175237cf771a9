"""pandaseq_b200 -- Python binding (ctypes) of libpandaseq_b200.so.

The product is the shared library (``include/pandaseq_b200.h``); this module is
the thin host-side mirror used by the tests and ``bench.py``.  PyTorch is used
only for device memory, streams and ``torch.distributed`` plumbing: every
compute call goes through the C ABI into the hand-written sm_100a kernels.

There is no fallback: if the library has not been built, importing the binding
symbols raises; if no GPU is present, compute calls raise ``PandaseqError``.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpandaseq_b200.so")

PB_MAX_LEN = 450
PB_PHREDMAX = 46
PB_NCOUNTERS = 16 + 2 * PB_MAX_LEN
ALGOS = {"simple_bayesian": 0, "pear": 1, "rdp_mle": 2, "flash": 3, "ea_util": 4, "stitch": 5, "uparse": 6}
STATUS = {0: "OK", 1: "BADR", 2: "NOFP", 3: "NORP", 4: "NOALGN", 5: "LOWQ", 6: "SKIP"}
FQ_ERRORS = {0: "OK", 1: "BADID", 2: "NOTPAIRED", 3: "EOF", 4: "BADNT", 5: "READLEN", 6: "BADSEQ", 7: "NOQUAL", 8: "LINELEN"}
TAG_PRESENT, TAG_ABSENT, TAG_OPTIONAL = 0, 1, 2
OUT_FASTA, OUT_FASTQ = 0, 1
C_COUNT, C_OK, C_LOWQ, C_NOALGN, C_BADR, C_NOFP, C_NORP, C_SLOW, C_LONGEST = range(9)
C_OVERLAPS = 16


class PandaseqError(RuntimeError):
    pass


class PbFilter(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ivalue", C.c_int32), ("dvalue", C.c_double), ("dvalue2", C.c_double), ("dvalue3", C.c_double)]


FILTERS = {"no_n": 1, "short": 2, "long": 3, "min_overlapbits": 4, "completely_miss_the_point": 5, "min_phred": 6, "pear_test": 7}
STATUS_FILTERED = 8
C_REJECTED = 9


class PbConfig(C.Structure):
    """mirror of pb_config (and, by construction, of oracle/panda_oracle.h po_config)"""
    _fields_ = [
        ("algo", C.c_int32), ("post_primers", C.c_int32),
        ("minoverlap", C.c_int64), ("maxoverlap", C.c_int64), ("num_kmers", C.c_int64),
        ("forward_trim", C.c_int64), ("reverse_trim", C.c_int64),
        ("forward_primer_length", C.c_int64), ("reverse_primer_length", C.c_int64),
        ("threshold", C.c_double), ("primer_penalty", C.c_double),
        ("sb_q", C.c_double), ("pear_random_base", C.c_double),
        ("forward_primer", C.c_char * PB_MAX_LEN), ("reverse_primer", C.c_char * PB_MAX_LEN),
        ("hang_forward_length", C.c_int64), ("hang_reverse_length", C.c_int64), ("hang_skip", C.c_int32), ("nfilters", C.c_int32),
        ("hang_threshold", C.c_double),
        ("hang_forward", C.c_char * PB_MAX_LEN), ("hang_reverse", C.c_char * PB_MAX_LEN),
        ("filters", PbFilter * 7),
    ]


NQ = PB_PHREDMAX + 1


class PbTables(C.Structure):
    _fields_ = [
        ("qual_nn", C.c_double),
        ("match_sb", C.c_double * NQ * NQ), ("mismatch_sb", C.c_double * NQ * NQ),
        ("match_pear", C.c_double * NQ * NQ), ("mismatch_pear", C.c_double * NQ * NQ),
        ("mismatch_rdp", C.c_double * NQ * NQ), ("mismatch_rdp_asm", C.c_double * NQ * NQ),
        ("match_uparse", C.c_double * NQ * NQ), ("mismatch_uparse", C.c_double * NQ * NQ),
        ("score", C.c_double * NQ), ("score_err", C.c_double * NQ),
    ]

    def as_dict(self):
        d = {"qual_nn": np.float64(self.qual_nn)}
        for name, _ in self._fields_[1:]:
            d[name] = np.ctypeslib.as_array(getattr(self, name)).copy()
        return d


PAIR_META_DTYPE = np.dtype([("off16", "<u4"), ("flen", "<u2"), ("rlen", "<u2")])
PAIR_RESULT_DTYPE = np.dtype([
    ("status", "u1"), ("slow", "u1"), ("overlap", "<u2"), ("seq_len", "<u2"), ("mismatches", "<u2"),
    ("degenerates", "<u2"), ("examined", "<u2"), ("fwd_offset", "<u2"), ("rev_offset", "<u2"),
    ("quality", "<f8"), ("est_prob", "<f8")])
assert PAIR_RESULT_DTYPE.itemsize == 32 and PAIR_META_DTYPE.itemsize == 8
SEQ_ID_DTYPE = np.dtype([("hdr_off", "<u4"), ("hdr_len", "<u2"), ("fmt", "u1"), ("reserved", "u1"),
                         ("inst_off", "<u2"), ("inst_len", "<u2"), ("run_off", "<u2"), ("run_len", "<u2"),
                         ("fc_off", "<u2"), ("fc_len", "<u2"), ("tag_off", "<u2"), ("tag_len", "<u2"),
                         ("lane", "<i4"), ("tile", "<i4"), ("x", "<i4"), ("y", "<i4"), ("sra", "<i4"), ("mate", "<i4")])
assert SEQ_ID_DTYPE.itemsize == 48
PANDA_SEQID_DTYPE = np.dtype([("instrument", "S100"), ("run", "S100"), ("flowcell", "S100"), ("lane", "<i4"), ("tile", "<i4"),
                              ("x", "<i4"), ("y", "<i4"), ("tag", "S50"), ("_pad", "V2")])      # pandaseq-common.h:235-247
assert PANDA_SEQID_DTYPE.itemsize == 368


class PbFastqInfo(C.Structure):
    _fields_ = [("records", C.c_uint64), ("limit", C.c_uint64), ("pairs", C.c_uint64), ("consumed_fwd", C.c_uint64),
                ("consumed_rev", C.c_uint64), ("error", C.c_int32), ("max_read_len", C.c_int32), ("stride16", C.c_uint32),
                ("reserved", C.c_uint32)]


class PbStreamInfo(C.Structure):
    _fields_ = [("records", C.c_uint64), ("pairs", C.c_uint64), ("consumed_fwd", C.c_uint64), ("consumed_rev", C.c_uint64),
                ("out_bytes", C.c_uint64), ("error", C.c_int32), ("reserved", C.c_int32)]


def make_config(algo="simple_bayesian", *, threshold=0.6, minoverlap=2, maxoverlap=0, forward_primer=None,
                reverse_primer=None, forward_trim=0, reverse_trim=0, primer_penalty=0.0, sb_q=0.36,
                pear_random_base=None, post_primers=False, num_kmers=2, hang_forward=None, hang_reverse=None, hang_skip=False,
                hang_threshold=None, filters=()) -> PbConfig:
    """Flat assembler configuration.  Primers are sequences of panda_nt codes *as the assembler
    stores them* (the reverse primer already complemented, args_assembler.c:222)."""
    cfg = PbConfig()
    cfg.algo = ALGOS[algo] if isinstance(algo, str) else int(algo)
    cfg.post_primers = int(post_primers)
    cfg.minoverlap, cfg.maxoverlap, cfg.num_kmers = int(minoverlap), int(maxoverlap), int(num_kmers)
    cfg.threshold = math.log(threshold)
    cfg.primer_penalty = float(primer_penalty)
    cfg.sb_q = float(sb_q)
    cfg.pear_random_base = math.log(0.25) if pear_random_base is None else float(pear_random_base)
    if forward_primer is not None and len(forward_primer):
        fp = bytes(int(x) & 0xFF for x in forward_primer)
        cfg.forward_primer_length = len(fp)
        C.memmove(C.byref(cfg, PbConfig.forward_primer.offset), fp, len(fp))
    else:
        cfg.forward_trim = int(forward_trim)
    if reverse_primer is not None and len(reverse_primer):
        rp = bytes(int(x) & 0xFF for x in reverse_primer)
        cfg.reverse_primer_length = len(rp)
        C.memmove(C.byref(cfg, PbConfig.reverse_primer.offset), rp, len(rp))
    else:
        cfg.reverse_trim = int(reverse_trim)
    # overhang trimmer: panda_nt codes as panda_trim_overhangs receives them (the reverse one already complemented)
    if hang_forward is not None and len(hang_forward):
        hf = bytes(int(x) & 0xFF for x in hang_forward)
        cfg.hang_forward_length = len(hf)
        C.memmove(C.byref(cfg, PbConfig.hang_forward.offset), hf, len(hf))
    if hang_reverse is not None and len(hang_reverse):
        hr = bytes(int(x) & 0xFF for x in hang_reverse)
        cfg.hang_reverse_length = len(hr)
        C.memmove(C.byref(cfg, PbConfig.hang_reverse.offset), hr, len(hr))
    cfg.hang_skip = int(bool(hang_skip))
    cfg.hang_threshold = cfg.threshold if hang_threshold is None else float(hang_threshold)
    # filters: sequence of (name, value) in module order, e.g. [("no_n", 0), ("short", 100), ("min_overlapbits", -20.0)]
    cfg.nfilters = len(filters)
    for k, (name, value) in enumerate(filters):
        cfg.filters[k].kind = FILTERS[name]
        if name == "min_overlapbits":
            cfg.filters[k].dvalue = float(value)
        elif name == "pear_test":              # (alpha, beta, cutoff), plugin_pear_test.c
            cfg.filters[k].dvalue, cfg.filters[k].dvalue2, cfg.filters[k].dvalue3 = (float(v) for v in value)
        else:
            cfg.filters[k].ivalue = int(value)
    return cfg


_lib = None


def lib() -> C.CDLL:
    """Load libpandaseq_b200.so (built in-tree by ``__graft_entry__.build()`` / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PandaseqError(f"{LIB_PATH} is missing: build it with `make -C pandaseq_b200/csrc` "
                            "(there is no Python or CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.pb_last_error.restype = C.c_char_p
    L.pb_device_count.restype = i32
    L.pb_context_create.argtypes = [i32, C.POINTER(vp)]
    L.pb_context_create.restype = i32
    L.pb_context_destroy.argtypes = [vp]
    L.pb_context_stream.argtypes = [vp]
    L.pb_context_stream.restype = vp
    L.pb_synchronize.argtypes = [vp]
    L.pb_synchronize.restype = i32
    L.pb_lanes_stats.argtypes = [vp, vp, vp]
    L.pb_lanes_stats.restype = i32
    L.pb_set_lanes.argtypes = [vp, i32]
    L.pb_set_lanes.restype = i32
    L.pb_set_timing.argtypes = [vp, i32]
    L.pb_set_timing.restype = i32
    L.pb_last_timing.argtypes = [vp, vp, vp]
    L.pb_last_timing.restype = i32
    L.pb_config_default.argtypes = [C.POINTER(PbConfig), i32]
    L.pb_get_tables.restype = C.POINTER(PbTables)
    L.pb_layout_host.argtypes = [sz, vp, vp, vp]
    L.pb_layout_host.restype = sz
    L.pb_pack_device.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp]
    L.pb_pack_device.restype = i32
    L.pb_assemble_device.argtypes = [vp, C.POINTER(PbConfig), sz, i32, vp, vp, vp, vp, vp, sz, vp]
    L.pb_assemble_device.restype = i32
    L.pb_assemble_host.argtypes = [vp, C.POINTER(PbConfig), sz, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.pb_assemble_host.restype = i32
    L.pb_assemble_host_codes.argtypes = [vp, C.POINTER(PbConfig), sz, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.pb_assemble_host_codes.restype = i32
    L.pb_assemble_host_packed.argtypes = [vp, C.POINTER(PbConfig), sz, i32, vp, vp, vp, vp, sz, vp]
    L.pb_assemble_host_packed.restype = i32
    L.pb_pack_host.argtypes = [sz, vp, vp, vp, vp, vp, vp]
    L.pb_pack_host.restype = None
    L.pb_posterior_table.argtypes = [C.POINTER(PbConfig), vp]
    L.pb_posterior_table.restype = i32
    L.pb_host_alloc.argtypes = [sz]
    L.pb_host_alloc.restype = vp
    L.pb_host_free.argtypes = [vp]
    L.pb_counters_merge.argtypes = [vp, vp]
    L.pb_fastq_parse_device.argtypes = [vp, vp, sz, vp, sz, i32, i32, sz, vp, sz, vp, vp, C.POINTER(PbFastqInfo)]
    L.pb_fastq_parse_device.restype = i32
    L.pb_format_device.argtypes = [vp, i32, sz, vp, vp, vp, sz, vp, vp, vp, sz, C.POINTER(sz)]
    L.pb_format_device.restype = i32
    L.pb_fastq_assemble_host.argtypes = [vp, C.POINTER(PbConfig), i32, i32, i32, vp, sz, vp, sz, i32, vp, sz, vp, C.POINTER(PbStreamInfo)]
    L.pb_fastq_assemble_host.restype = i32
    L.pb_seq_id_expand.argtypes = [vp, vp, vp]
    L.panda_max_len.restype = sz
    _lib = L
    return L


def _check(status: int, what: str):
    if status != 0:
        raise PandaseqError(f"{what}: status {status}: {lib().pb_last_error().decode(errors='replace')}")


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def unpack_nt(packed: np.ndarray) -> np.ndarray:
    """(n, stride/2) packed merged reads (4 bit per base, low nibble first) -> (n, stride) panda_nt codes"""
    out = np.empty((packed.shape[0], packed.shape[1] * 2), dtype=np.uint8)
    out[:, 0::2] = packed & 0x0F
    out[:, 1::2] = packed >> 4
    return out


def record_bytes(flen, rlen):
    flen, rlen = np.asarray(flen, dtype=np.int64), np.asarray(rlen, dtype=np.int64)
    b = ((flen + 7) // 8) * 4 + ((rlen + 7) // 8) * 4 + ((flen + 3) // 4) * 4 + ((rlen + 3) // 4) * 4
    return (b + 15) & ~15


class Context:
    """One device context (stream + LUT block + staging) on a GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().pb_context_create(int(device), C.byref(self._h)), "pb_context_create")
        self.device = int(device)

    def close(self):
        if self._h:
            lib().pb_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream_handle(self) -> int:
        return int(lib().pb_context_stream(self._h))

    def synchronize(self):
        _check(lib().pb_synchronize(self._h), "pb_synchronize")

    def lanes_stats(self):
        """(pairs launched on the lane-per-pair kernel, pairs it handed to the general kernel) since the context was created."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(lib().pb_lanes_stats(self._h, C.byref(a), C.byref(b)), "pb_lanes_stats")
        return int(a.value), int(b.value)

    def set_lanes(self, mode: int):
        """0 = general kernel only, 1 = two-kernel path where it applies (seeding by the diagonal sweep), 2 = two-kernel path with
        the hash-join seeding kernel, -1 = follow PANDASEQ_B200_LANES / PANDASEQ_B200_SWEEP."""
        _check(lib().pb_set_lanes(self._h, int(mode)), "pb_set_lanes")

    def set_timing(self, on: bool):
        _check(lib().pb_set_timing(self._h, 1 if on else 0), "pb_set_timing")

    def last_timing(self):
        """(kind, [seed ms, lanes ms, general ms]) of the last assemble_device call made with timing on."""
        kind = C.c_int(0)
        ms = (C.c_float * 3)()
        _check(lib().pb_last_timing(self._h, C.byref(kind), ms), "pb_last_timing")
        return int(kind.value), [float(x) for x in ms]

    # ---- host-buffer (e2e) path ----------------------------------------------------
    def assemble_host(self, cfg: PbConfig, batch, *, want_nt=True, want_p=False, seq_stride=None):
        """batch: synth.FlatBatch.  Returns dict(results, seq_nt, seq_p, counters) of numpy arrays."""
        n = batch.n
        if seq_stride is None:
            fl, rl = batch.lengths()
            seq_stride = int((fl + rl).max()) if n else 0
            seq_stride = (seq_stride + 15) & ~15
        res = np.zeros(n, dtype=PAIR_RESULT_DTYPE)
        nt = np.zeros((n, seq_stride // 2), dtype=np.uint8) if want_nt else None
        p = np.zeros((n, seq_stride), dtype=np.float64) if want_p else None
        counters = np.zeros(PB_NCOUNTERS, dtype=np.int64)
        f_data, r_data = np.ascontiguousarray(batch.f_data), np.ascontiguousarray(batch.r_data)
        f_off, r_off = np.ascontiguousarray(batch.f_off, dtype=np.uint64), np.ascontiguousarray(batch.r_off, dtype=np.uint64)
        _check(lib().pb_assemble_host(self._h, C.byref(cfg), n, _np_ptr(f_data), _np_ptr(f_off), _np_ptr(r_data), _np_ptr(r_off),
                                      _np_ptr(res), _np_ptr(nt), _np_ptr(p), seq_stride, _np_ptr(counters)), "pb_assemble_host")
        return dict(results=res, seq_nt=unpack_nt(nt) if nt is not None else None, seq_nt_packed=nt, seq_p=p, counters=counters,
                    seq_stride=seq_stride)

    def assemble_host_codes(self, cfg: PbConfig, batch, *, seq_stride=None):
        """As assemble_host(want_p=True), the per-base log p shipped as 16-bit codes (pb_assemble_host_codes) and expanded here
        with pb_posterior_table(): dict(results, seq_nt, seq_p, seq_code, counters)."""
        n = batch.n
        if seq_stride is None:
            fl, rl = batch.lengths()
            seq_stride = ((int((fl + rl).max()) if n else 0) + 15) & ~15
        res = np.zeros(n, dtype=PAIR_RESULT_DTYPE)
        nt = np.zeros((n, seq_stride // 2), dtype=np.uint8)
        code = np.zeros((n, seq_stride), dtype=np.uint16)
        counters = np.zeros(PB_NCOUNTERS, dtype=np.int64)
        f_data, r_data = np.ascontiguousarray(batch.f_data), np.ascontiguousarray(batch.r_data)
        f_off, r_off = np.ascontiguousarray(batch.f_off, dtype=np.uint64), np.ascontiguousarray(batch.r_off, dtype=np.uint64)
        _check(lib().pb_assemble_host_codes(self._h, C.byref(cfg), n, _np_ptr(f_data), _np_ptr(f_off), _np_ptr(r_data), _np_ptr(r_off),
                                            _np_ptr(res), _np_ptr(nt), _np_ptr(code), seq_stride, _np_ptr(counters)), "pb_assemble_host_codes")
        table = posterior_table(cfg)
        p = table[np.minimum(code, len(table) - 1)]
        p[np.arange(seq_stride)[None, :] >= res["seq_len"][:, None]] = 0.0
        return dict(results=res, seq_nt=unpack_nt(nt), seq_nt_packed=nt, seq_p=p, seq_code=code, counters=counters, seq_stride=seq_stride)

    def assemble_host_packed(self, cfg: PbConfig, reads, meta, max_len, *, seq_stride):
        """reads / meta: numpy arrays in the packed layout (pack_host); -> dict(results, seq_nt, counters)."""
        n = len(meta)
        res = np.zeros(n, dtype=PAIR_RESULT_DTYPE)
        nt = np.zeros((n, seq_stride // 2), dtype=np.uint8)
        counters = np.zeros(PB_NCOUNTERS, dtype=np.int64)
        _check(lib().pb_assemble_host_packed(self._h, C.byref(cfg), n, int(max_len), _np_ptr(reads), _np_ptr(meta), _np_ptr(res), _np_ptr(nt),
                                             seq_stride, _np_ptr(counters)), "pb_assemble_host_packed")
        return dict(results=res, seq_nt=unpack_nt(nt), seq_nt_packed=nt, seq_p=None, counters=counters, seq_stride=seq_stride)

    # ---- device-resident path (torch tensors own the HBM) ---------------------------
    def pack_device(self, f_data, f_off, r_data, r_off):
        """flat AoS torch CUDA tensors -> (reads u8, meta (n,2) i32 view, max_len, record bytes).
        The record offsets (an exclusive scan of the record sizes) are computed with torch on the device."""
        import torch
        n = f_off.numel() - 1
        flen, rlen = f_off[1:] - f_off[:-1], r_off[1:] - r_off[:-1]
        rb = ((flen + 7) // 8) * 4 + ((rlen + 7) // 8) * 4 + ((flen + 3) // 4) * 4 + ((rlen + 3) // 4) * 4
        rb16 = (rb + 15) // 16
        rec_off = torch.zeros(n + 1, dtype=torch.int64, device=f_data.device)
        rec_off[1:] = torch.cumsum(rb16, 0)
        total = int(rec_off[-1].item()) * 16
        rec_off32 = rec_off[:-1].to(torch.int32).contiguous()   # < 2^31 units of 16 B
        reads = torch.empty(total + 16, dtype=torch.uint8, device=f_data.device)
        meta = torch.empty((n, 2), dtype=torch.int32, device=f_data.device)
        max_len = int(torch.maximum(flen.max(), rlen.max()).item()) if n else 0
        self._sync_from_torch()
        _check(lib().pb_pack_device(self._h, n, f_data.data_ptr(), f_off.data_ptr(), r_data.data_ptr(), r_off.data_ptr(),
                                    rec_off32.data_ptr(), reads.data_ptr(), meta.data_ptr()), "pb_pack_device")
        self.synchronize()
        return reads, meta, max_len, total

    def _sync_from_torch(self):
        import torch
        torch.cuda.current_stream(self.device).synchronize()

    def assemble_device(self, cfg: PbConfig, n, max_len, reads, meta, results, seq_nt, seq_p, seq_stride, counters):
        """Asynchronous launch on the context's stream; all tensors are torch CUDA tensors
        (results: (n,32) uint8, counters: (PB_NCOUNTERS,) int64)."""
        _check(lib().pb_assemble_device(self._h, C.byref(cfg), int(n), int(max_len), reads.data_ptr(), meta.data_ptr(),
                                        results.data_ptr(), seq_nt.data_ptr() if seq_nt is not None else None,
                                        seq_p.data_ptr() if seq_p is not None else None, int(seq_stride), counters.data_ptr()),
               "pb_assemble_device")


    # ---- FASTQ text in, FASTA/FASTQ text out -------------------------------------------------------
    def fastq_parse_device(self, fwd, rev, *, qualmin=33, policy=TAG_PRESENT, max_records=None):
        """fwd/rev: uint8 torch CUDA tensors holding FASTQ text.  Returns dict(info, reads, meta, ids) -- torch CUDA
        tensors (meta: (max_records, 2) int32 view of pb_pair_meta; ids: (max_records, 48) uint8) and the info dict."""
        import torch
        dev = fwd.device
        if max_records is None:
            max_records = int(min(fwd.numel(), rev.numel()) // 7 + 1)       # "@\n\n+\n\n" is the shortest record
        cap = max(max_records, 1) * int(record_bytes(PB_MAX_LEN, PB_MAX_LEN))
        reads = torch.zeros(cap, dtype=torch.uint8, device=dev)
        meta = torch.zeros((max(max_records, 1), 2), dtype=torch.int32, device=dev)
        ids = torch.zeros((max(max_records, 1), 48), dtype=torch.uint8, device=dev)
        info = PbFastqInfo()
        self._sync_from_torch()
        _check(lib().pb_fastq_parse_device(self._h, fwd.data_ptr(), fwd.numel(), rev.data_ptr(), rev.numel(), int(qualmin), int(policy),
                                           int(max_records), reads.data_ptr(), reads.numel(), meta.data_ptr(), ids.data_ptr(),
                                           C.byref(info)), "pb_fastq_parse_device")
        d = {k: getattr(info, k) for k, _ in PbFastqInfo._fields_}
        return dict(info=d, reads=reads, meta=meta, ids=ids)

    def format_device(self, fmt, n, results, seq_nt, seq_p, seq_stride, ids, fwd, capacity=None):
        """-> bytes of the FASTA/FASTQ text of records [0, n) (torch CUDA tensors in, text copied back)."""
        import torch
        self._sync_from_torch()
        total = C.c_size_t(0)
        args = lambda t, cap: (self._h, int(fmt), int(n), results.data_ptr(), seq_nt.data_ptr(), seq_p.data_ptr() if seq_p is not None else None,
                               int(seq_stride), ids.data_ptr(), fwd.data_ptr(), t, cap, C.byref(total))
        _check(lib().pb_format_device(*args(None, 0)), "pb_format_device (length)")
        text = torch.zeros(max(int(total.value), 1), dtype=torch.uint8, device=fwd.device)
        _check(lib().pb_format_device(*args(text.data_ptr(), text.numel())), "pb_format_device")
        return bytes(text[:int(total.value)].cpu().numpy())

    def fastq_assemble_host(self, cfg: PbConfig, fwd, rev, *, qualmin=33, policy=TAG_PRESENT, out_format=OUT_FASTA, final=True,
                            out=None, out_capacity=None):
        """fwd/rev: bytes or uint8 numpy arrays / pinned torch tensors with FASTQ text.  Returns (text bytes or None, info dict,
        counters).  `out`: optional preallocated uint8 buffer (numpy or pinned torch) receiving the text."""
        def ptr_len(x):
            if isinstance(x, (bytes, bytearray)):
                a = np.frombuffer(x, dtype=np.uint8)
                return a, a.ctypes.data, a.size
            if isinstance(x, np.ndarray):
                return x, x.ctypes.data, x.size
            return x, x.data_ptr(), x.numel()
        fk, fp, fl = ptr_len(fwd)
        rk, rp, rl = ptr_len(rev)
        own = out is None
        if own:
            out = np.zeros(int(out_capacity if out_capacity is not None else fl + rl + 1024), dtype=np.uint8)
        ok, op, ol = ptr_len(out)
        counters = np.zeros(PB_NCOUNTERS, dtype=np.int64)
        info = PbStreamInfo()
        _check(lib().pb_fastq_assemble_host(self._h, C.byref(cfg), int(qualmin), int(policy), int(out_format), fp, fl, rp, rl, int(bool(final)),
                                            op, ol, _np_ptr(counters), C.byref(info)), "pb_fastq_assemble_host")
        d = {k: getattr(info, k) for k, _ in PbStreamInfo._fields_}
        text = bytes(out[:d["out_bytes"]]) if own else None
        return text, d, counters


PAIR_META_DTYPE = np.dtype([("off16", "<u4"), ("flen", "<u2"), ("rlen", "<u2")])


def pack_host(batch):
    """FlatBatch -> (reads uint8, meta PAIR_META_DTYPE, longest read): the packed layout written on the host (pb_pack_host)."""
    n = batch.n
    f_data, r_data = np.ascontiguousarray(batch.f_data), np.ascontiguousarray(batch.r_data)
    f_off, r_off = np.ascontiguousarray(batch.f_off, dtype=np.uint64), np.ascontiguousarray(batch.r_off, dtype=np.uint64)
    rec = np.zeros(max(n, 1), dtype=np.uint32)
    total = int(lib().pb_layout_host(n, _np_ptr(f_off), _np_ptr(r_off), _np_ptr(rec)))
    reads = np.zeros(total + 16, dtype=np.uint8)
    meta = np.zeros(n, dtype=PAIR_META_DTYPE)
    lib().pb_pack_host(n, _np_ptr(f_data), _np_ptr(f_off), _np_ptr(r_data), _np_ptr(r_off), _np_ptr(reads), _np_ptr(meta))
    fl, rl = batch.lengths()
    return reads, meta, int(max(fl.max(), rl.max())) if n else 0


def posterior_table(cfg: PbConfig) -> np.ndarray:
    """the 2 x 48 x 48 table the per-base codes of pb_assemble_host_codes index (pb_posterior_table)"""
    t = np.zeros(2 * 48 * 48, dtype=np.float64)
    _check(lib().pb_posterior_table(C.byref(cfg), _np_ptr(t)), "pb_posterior_table")
    return t


def expand_ids(ids: np.ndarray, fwd_text: bytes) -> np.ndarray:
    """pb_seq_id records (SEQ_ID_DTYPE) -> panda_seq_identifier records (PANDA_SEQID_DTYPE), via pb_seq_id_expand"""
    ids = np.ascontiguousarray(ids)
    out = np.zeros(len(ids), dtype=PANDA_SEQID_DTYPE)
    buf = C.create_string_buffer(fwd_text, len(fwd_text) + 1)
    for i in range(len(ids)):
        lib().pb_seq_id_expand(ids[i:i + 1].ctypes.data, C.cast(buf, C.c_void_p), out[i:i + 1].ctypes.data)
    return out


from . import synth  # noqa: E402,F401  (torch is imported lazily by callers that generate data)


def tables() -> dict:
    return lib().pb_get_tables().contents.as_dict()
