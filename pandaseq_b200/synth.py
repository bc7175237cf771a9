"""Deterministic synthetic read pairs.

Every pair is a random template of length L; the forward read is its first RL_f
bases, the reverse read the last RL_r bases of the opposite strand.  Reads are
returned the way the reference's API sees them (pandaseq-common.h:210-219,
fastq.c:154): ``panda_qual`` arrays of (nt, qual) bytes with nt the 4-bit
one-hot code, the reverse read complemented but in read order.

Quality model: PHRED for cycle i is clamp(round(N(36 - 12*i/RL, 4)), 2, 41); a
base is substituted by a different one with probability 10^(-q/10).  Optionally
0.1 % of bases become N (q = 2) and 5 % of reads get a trailing run of '#'
(q = 2), which exercises the N and B-cliff paths of align().

Runs on whatever torch device is asked for (CPU here, CUDA on the GPU box); the
stream of random numbers differs between device types, the distribution does not.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

# BASELINE.json configs -> generator parameters (cfg 4's primers: pandaseq.1:354)
FWD_PRIMER = "CCTACGGGAGGCAGCAG"
REV_PRIMER = "ATTACCGCGGCTGCTGG"
CONFIGS = {
    1: dict(n=10_000, rl=(150, 150), tmpl=(180, 280), seed=1, algo="simple_bayesian"),
    2: dict(n=10_000_000, rl=(150, 150), tmpl=(180, 280), seed=2, algo="simple_bayesian"),
    3: dict(n=10_000_000, rl=(250, 250), tmpl=(300, 450), seed=3, algo="pear"),
    4: dict(n=10_000_000, rl=(300, 300), tmpl=(350, 500), seed=4, algo="rdp_mle", primers=True),
    5: dict(n=100_000_000, rl=(75, 300), tmpl=None, seed=5, algo="simple_bayesian", mixed=True),
}

_NT = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def encode(seq: str) -> np.ndarray:
    return np.array([_NT[c] for c in seq.upper()], dtype=np.uint8)


def revcomp(seq: str) -> str:
    return "".join(_COMP[c] for c in reversed(seq.upper()))


@dataclass
class FlatBatch:
    """n pairs as flat ``panda_qual`` arrays: data[k] = (nt, qual), off has n+1 element offsets."""

    f_data: np.ndarray  # (Nf, 2) uint8
    f_off: np.ndarray   # (n+1,) uint64
    r_data: np.ndarray
    r_off: np.ndarray

    @property
    def n(self) -> int:
        return len(self.f_off) - 1

    def lengths(self):
        return np.diff(self.f_off).astype(np.int64), np.diff(self.r_off).astype(np.int64)

    def slice(self, a: int, b: int) -> "FlatBatch":
        fo, ro = self.f_off[a:b + 1], self.r_off[a:b + 1]
        return FlatBatch(self.f_data[int(fo[0]):int(fo[-1])], (fo - fo[0]).astype(np.uint64),
                         self.r_data[int(ro[0]):int(ro[-1])], (ro - ro[0]).astype(np.uint64))

    def pair(self, i: int):
        f = self.f_data[int(self.f_off[i]):int(self.f_off[i + 1])]
        r = self.r_data[int(self.r_off[i]):int(self.r_off[i + 1])]
        return f, r

    @staticmethod
    def from_pairs(pairs) -> "FlatBatch":
        """pairs: iterable of (f_nt, f_q, r_nt, r_q) 1-D integer arrays (read order)."""
        fs, rs, fo, ro = [], [], [0], [0]
        for f_nt, f_q, r_nt, r_q in pairs:
            f = np.stack([np.asarray(f_nt, dtype=np.int64) & 0xFF, np.asarray(f_q, dtype=np.int64) & 0xFF], axis=1).astype(np.uint8).reshape(-1, 2)
            r = np.stack([np.asarray(r_nt, dtype=np.int64) & 0xFF, np.asarray(r_q, dtype=np.int64) & 0xFF], axis=1).astype(np.uint8).reshape(-1, 2)
            fs.append(f)
            rs.append(r)
            fo.append(fo[-1] + len(f))
            ro.append(ro[-1] + len(r))
        cat = lambda xs: np.concatenate(xs, axis=0) if xs else np.zeros((0, 2), np.uint8)
        return FlatBatch(cat(fs), np.array(fo, dtype=np.uint64), cat(rs), np.array(ro, dtype=np.uint64))

    @staticmethod
    def concat(batches) -> "FlatBatch":
        batches = list(batches)
        f = np.concatenate([b.f_data for b in batches], axis=0)
        r = np.concatenate([b.r_data for b in batches], axis=0)
        fo, ro = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)]
        fb = rb = 0
        for b in batches:
            fo.append(b.f_off[1:] + np.uint64(fb))
            ro.append(b.r_off[1:] + np.uint64(rb))
            fb += len(b.f_data)
            rb += len(b.r_data)
        return FlatBatch(f, np.concatenate(fo), r, np.concatenate(ro))


@dataclass
class RectBatch:
    """Torch tensors on one device: (n, RLmax) nt/qual per read plus per-pair lengths."""

    f_nt: torch.Tensor  # uint8
    f_q: torch.Tensor   # uint8
    r_nt: torch.Tensor
    r_q: torch.Tensor
    flen: torch.Tensor  # int64
    rlen: torch.Tensor

    def to_flat_tensors(self):
        """-> (f_data (Nf,2) u8, f_off (n+1) i64, r_data, r_off) on the same device."""
        out = []
        for nt, q, ln in ((self.f_nt, self.f_q, self.flen), (self.r_nt, self.r_q, self.rlen)):
            n, w = nt.shape
            off = torch.zeros(n + 1, dtype=torch.int64, device=nt.device)
            off[1:] = torch.cumsum(ln, 0)
            if bool((ln == w).all()):
                data = torch.stack([nt, q], dim=2).reshape(-1, 2)
            else:
                keep = torch.arange(w, device=nt.device)[None, :] < ln[:, None]
                data = torch.stack([nt[keep], q[keep]], dim=1)
            out += [data.contiguous(), off]
        return tuple(out)

    def to_flat(self) -> FlatBatch:
        f, fo, r, ro = self.to_flat_tensors()
        return FlatBatch(f.cpu().numpy(), fo.cpu().numpy().astype(np.uint64), r.cpu().numpy(), ro.cpu().numpy().astype(np.uint64))


def _quals(gen, n, rl, lens, device):
    i = torch.arange(rl, device=device, dtype=torch.float32)[None, :]
    mu = 36.0 - 12.0 * i / lens[:, None].to(torch.float32)
    q = torch.round(mu + 4.0 * torch.randn((n, rl), generator=gen, device=device))
    return q.clamp_(2, 41).to(torch.int64)


def _observe(gen, true_code, q, device, n_rate):
    """apply substitution errors and optional Ns; returns (nt, q) uint8"""
    n, rl = true_code.shape
    perr = torch.pow(10.0, -q.to(torch.float32) / 10.0)
    err = torch.rand((n, rl), generator=gen, device=device) < perr
    shift = torch.randint(1, 4, (n, rl), generator=gen, device=device)
    code = torch.where(err, (true_code + shift) & 3, true_code)
    nt = (1 << code).to(torch.uint8)
    if n_rate > 0:
        isn = torch.rand((n, rl), generator=gen, device=device) < n_rate
        nt = torch.where(isn, torch.full_like(nt, 15), nt)
        q = torch.where(isn, torch.full_like(q, 2), q)
    return nt, q


def _btail(gen, q, lens, device, rate):
    if rate <= 0:
        return q
    n, rl = q.shape
    has = torch.rand((n,), generator=gen, device=device) < rate
    run = torch.randint(1, 31, (n,), generator=gen, device=device)
    start = torch.where(has, lens - run, lens)
    pos = torch.arange(rl, device=device)[None, :]
    return torch.where(pos >= start[:, None], torch.full_like(q, 2), q)


def generate(n: int, rl=(150, 150), tmpl=(180, 280), seed=1, device="cpu", mixed=False, primers=False,
             n_rate=0.0, btail_rate=0.0, chunk_index=0) -> RectBatch:
    """One chunk of n synthetic pairs.  (seed, chunk_index) fixes the chunk."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1_000_003 + int(chunk_index))
    if mixed:
        lo, hi = rl
        flen = torch.randint(lo, hi + 1, (n,), generator=gen, device=device)
        rlen = torch.randint(lo, hi + 1, (n,), generator=gen, device=device)
        longest = torch.maximum(flen, rlen)
        tlo, thi = longest + 10, flen + rlen - 10
        thi = torch.maximum(thi, tlo)
        L = tlo + (torch.rand((n,), generator=gen, device=device) * (thi - tlo + 1).to(torch.float32)).to(torch.int64)
        L = torch.minimum(L, thi)
        rlf = rlr = hi
    else:
        rlf, rlr = rl
        flen = torch.full((n,), rlf, dtype=torch.int64, device=device)
        rlen = torch.full((n,), rlr, dtype=torch.int64, device=device)
        L = torch.randint(tmpl[0], tmpl[1] + 1, (n,), generator=gen, device=device)
    fp = rp = None
    if primers:
        fp = torch.from_numpy(np.log2(encode(FWD_PRIMER)).astype(np.int64)).to(device)
        rp = torch.from_numpy(np.log2(encode(revcomp(REV_PRIMER))).astype(np.int64)).to(device)
        L = L + len(fp) + len(rp)
    Lmax = int(L.max().item())
    # Extended template: the insert occupies [rlr, rlr+L); what lies outside is random sequence, so an
    # insert shorter than a read ("read-through") simply runs on into unrelated bases.
    width = rlr + max(Lmax, rlf) + rlf
    tcode = torch.randint(0, 4, (n, width), generator=gen, device=device)
    if primers:
        tcode[:, rlr:rlr + len(fp)] = fp[None, :]
        idx = (rlr + L - len(rp))[:, None] + torch.arange(len(rp), device=device)[None, :]
        tcode.scatter_(1, idx, rp[None, :].expand(n, -1))
    # forward read i: insert[i]; reverse read k: insert[L-1-k] (complemented, read order)
    fpos = (rlr + torch.arange(rlf, device=device))[None, :].expand(n, -1)
    f_true = torch.gather(tcode, 1, fpos)
    rpos = rlr + L[:, None] - 1 - torch.arange(rlr, device=device)[None, :]
    r_true = torch.gather(tcode, 1, rpos)
    f_q = _btail(gen, _quals(gen, n, rlf, flen, device), flen, device, btail_rate)
    r_q = _btail(gen, _quals(gen, n, rlr, rlen, device), rlen, device, btail_rate)
    f_nt, f_q = _observe(gen, f_true, f_q, device, n_rate)
    r_nt, r_q = _observe(gen, r_true, r_q, device, n_rate)
    return RectBatch(f_nt, f_q.to(torch.uint8), r_nt, r_q.to(torch.uint8), flen, rlen)


def generate_config(cfg_id: int, n: int | None = None, device="cpu", n_rate=0.0, btail_rate=0.0, chunk_index=0) -> RectBatch:
    c = CONFIGS[cfg_id]
    return generate(n if n is not None else c["n"], rl=c["rl"], tmpl=c.get("tmpl") or (0, 0), seed=c["seed"], device=device,
                    mixed=c.get("mixed", False), primers=c.get("primers", False), n_rate=n_rate, btail_rate=btail_rate,
                    chunk_index=chunk_index)


# ---- FASTQ text of a batch (what fastq.c parses; SURVEY.md §8d header format) -------------------------
_LETTERS = np.frombuffer(b"NACMGRSVTWYHKDBN", dtype=np.uint8)        # nt.c:25
_COMP4 = np.array([((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3) for c in range(16)], dtype=np.uint8)


def fastq_text(data, off, mate: int, *, complement: bool, qual_offset=33, seed=7, crlf=False):
    """FASTQ text (uint8 torch tensor, same device as `data`) of flat reads: data (N,2) uint8 (nt, qual), off (n+1).
    Headers are CASAVA 1.7, fixed width: ``@M01271:10:000000000-A3WGH:1:<tile 4d>:<x 5d>:<y 5d> <mate>:N:0:1``; tile/x/y
    are a deterministic function of (seed, pair index), identical for both mates.  complement=True writes the letter of the
    complementary base, i.e. what the sequencer reports for a reverse read that the API holds complemented (fastq.c:154)."""
    data = torch.as_tensor(data)
    dev = data.device
    off = torch.as_tensor(np.asarray(off).astype(np.int64) if not torch.is_tensor(off) else off).to(dev).to(torch.int64)
    n = off.numel() - 1
    lens = off[1:] - off[:-1]
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    h = (idx * 2654435761 + seed * 40503) & 0x7FFFFFFF
    tile = 1101 + (h % 1019)
    x = 10000 + ((h // 1019) % 20000)
    y = 10000 + ((idx * 7919 + seed) % 20000)
    prefix = torch.tensor(list(b"@M01271:10:000000000-A3WGH:1:"), dtype=torch.uint8, device=dev)
    suffix = torch.tensor(list(b" %d:N:0:1" % mate), dtype=torch.uint8, device=dev)

    def digits(v, w):
        return torch.stack([(v // (10 ** (w - 1 - k))) % 10 + 48 for k in range(w)], dim=1).to(torch.uint8)

    colon = torch.full((n, 1), 58, dtype=torch.uint8, device=dev)
    hdr = torch.cat([prefix[None, :].expand(n, -1), digits(tile, 4), colon, digits(x, 5), colon, digits(y, 5),
                     suffix[None, :].expand(n, -1)], dim=1)
    hw = hdr.shape[1]
    eol = 2 if crlf else 1
    rec_len = hw + eol + lens + eol + 1 + eol + lens + eol
    rec_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rec_off[1:] = torch.cumsum(rec_len, 0)
    total = int(rec_off[-1].item())
    out = torch.full((total,), 10, dtype=torch.uint8, device=dev)          # every byte not written below is '\n'
    start = rec_off[:-1]
    pos = start[:, None] + torch.arange(hw, device=dev)[None, :]
    out[pos.reshape(-1)] = hdr.reshape(-1)
    nt = data[:, 0].to(torch.int64) & 15
    if complement:
        nt = torch.from_numpy(_COMP4).to(dev)[nt].to(torch.int64)
    letters = torch.from_numpy(_LETTERS.copy()).to(dev)[nt]
    quals = (data[:, 1].to(torch.int16) + qual_offset).to(torch.uint8)
    rec = torch.repeat_interleave(idx, lens)
    k = torch.arange(data.shape[0], device=dev, dtype=torch.int64) - off[:-1][rec]
    seq_pos = start[rec] + hw + eol + k
    out[seq_pos] = letters
    plus_pos = start + hw + eol + lens + eol
    out[plus_pos] = 43
    out[seq_pos + lens[rec] + eol + 1 + eol] = quals
    if crlf:
        for p in (start + hw, start + hw + eol + lens, plus_pos + 1, start + rec_len - eol):
            out[p] = 13
    return out


def fastq_pair(batch: "FlatBatch", device="cpu", **kw):
    """(forward text, reverse text) uint8 tensors for a FlatBatch"""
    f = fastq_text(torch.from_numpy(np.ascontiguousarray(batch.f_data)).to(device), batch.f_off, 1, complement=False, **kw)
    r = fastq_text(torch.from_numpy(np.ascontiguousarray(batch.r_data)).to(device), batch.r_off, 2, complement=True, **kw)
    return f, r
