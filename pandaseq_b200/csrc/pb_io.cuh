/* pb_io.cuh -- sm_100a kernels for the two stages either side of the assembly kernel.
 *
 *   FASTQ text -> packed records (the reference's fastq.c:44-193 reader over linebuf.c:57-89 lines, identifiers by
 *   seqid.c:136-285, letters by nt.c:48-124):
 *     nl_count / nl_scan / nl_write   line index: position of every '\n' of both texts (HBM-bound byte scan)
 *     fq_geometry                     record count, bytes consumed, longest read, record stride
 *     fq_ids                          one THREAD per record: both header lines parsed, compared (fastq.c:121-139)
 *     fq_reads                        one WARP per record: letters -> 4-bit codes (reverse read complemented and laid
 *                                     out in template order), quality characters -> PHRED with fastq.c:44's clamp,
 *                                     the '+' line and length checks, straight into the assemble kernel's layout
 *     fq_finish                       first failing record (the reader stops there), pairs delivered
 *   assembled pairs -> FASTA/FASTQ text (output.c:85-126):
 *     fmt_length / (scan) / fmt_write
 *
 * All of it is byte/integer work bounded by HBM bandwidth; nothing here is a contraction.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pb_internal.h"

namespace pbio {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int NL_THREADS = 256;
constexpr int NL_BYTES_PER_THREAD = 32;
constexpr int NL_TILE = NL_THREADS * NL_BYTES_PER_THREAD;    /* 8 KB of text per CTA */
constexpr uint64_t NO_ERROR_KEY = ~0ull;

struct TextView {
	const uint8_t *text;
	unsigned long long bytes;
	uint32_t *nl;               /* newline positions */
	unsigned nl_cap;
	uint32_t *block_cnt;        /* per-tile newline counts, then their exclusive scan */
	unsigned nblocks;
};

/* Device-resident state of one parse (written by the kernels, copied back once). */
struct ParseState {
	unsigned nl_total[2];       /* newlines in the forward / reverse text */
	unsigned records;
	unsigned max_len[2];        /* longest sequence line of the records, clamped to PB_MAX_LEN */
	unsigned stride16;
	unsigned long long consumed[2];
	unsigned long long err_key; /* min over failing records of (record << 8 | stage << 4 | code) */
	unsigned long long limit;
	unsigned long long pairs;
	int error;
	int nl_overflow;
};

/* 32 bytes of text -> bit b set iff byte b is '\n' (bytes past `valid` ignored) */
__device__ __forceinline__ unsigned newline_mask(const uint8_t *p, long long valid) {
	unsigned mask = 0;
	if (valid >= 32) {
		const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
		const unsigned w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const unsigned eq = __vcmpeq4(w[k], 0x0A0A0A0Au) & 0x01010101u;        /* bit 8j set iff byte j matches */
			mask |= (((eq * 0x01020408u) >> 24) & 15u) << (4 * k);             /* gather the four flags: byte j -> bit j */
		}
	} else {
		for (int k = 0; k < valid; k++)
			if (p[k] == '\n')
				mask |= 1u << k;
	}
	return mask;
}

__global__ void __launch_bounds__(NL_THREADS) nl_count(TextView tf, TextView tr) {
	const TextView &t = blockIdx.y ? tr : tf;
	if (blockIdx.x >= t.nblocks)
		return;
	const long long base = (long long) blockIdx.x * NL_TILE + (long long) threadIdx.x * NL_BYTES_PER_THREAD;
	const long long valid = (long long) t.bytes - base;
	int c = valid > 0 ? __popc(newline_mask(t.text + base, valid)) : 0;
	c = __reduce_add_sync(FULL, c);
	__shared__ int wsum[NL_THREADS / 32];
	if ((threadIdx.x & 31) == 0)
		wsum[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		int s = 0;
		for (int k = 0; k < NL_THREADS / 32; k++)
			s += wsum[k];
		t.block_cnt[blockIdx.x] = (uint32_t) s;
	}
}

/* exclusive scan of block_cnt in place, one CTA per text; total -> st->nl_total[] */
__global__ void __launch_bounds__(1024) nl_scan(TextView tf, TextView tr, ParseState *st) {
	const TextView &t = blockIdx.x ? tr : tf;
	__shared__ unsigned part[1024];
	const unsigned per = (t.nblocks + 1023) / 1024;
	const unsigned lo = min(threadIdx.x * per, t.nblocks), hi = min(lo + per, t.nblocks);
	unsigned s = 0;
	for (unsigned k = lo; k < hi; k++)
		s += t.block_cnt[k];
	part[threadIdx.x] = s;
	__syncthreads();
	for (int d = 1; d < 1024; d <<= 1) {             /* Hillis-Steele over the 1024 partial sums */
		unsigned v = threadIdx.x >= (unsigned) d ? part[threadIdx.x - d] : 0u;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	unsigned run = part[threadIdx.x] - s;
	for (unsigned k = lo; k < hi; k++) {
		const unsigned c = t.block_cnt[k];
		t.block_cnt[k] = run;
		run += c;
	}
	if (threadIdx.x == 1023) {
		st->nl_total[blockIdx.x] = part[1023];
		if (part[1023] > t.nl_cap)
			st->nl_overflow = 1;
	}
}

__global__ void __launch_bounds__(NL_THREADS) nl_write(TextView tf, TextView tr) {
	const TextView &t = blockIdx.y ? tr : tf;
	if (blockIdx.x >= t.nblocks)
		return;
	const long long base = (long long) blockIdx.x * NL_TILE + (long long) threadIdx.x * NL_BYTES_PER_THREAD;
	const long long valid = (long long) t.bytes - base;
	unsigned mask = valid > 0 ? newline_mask(t.text + base, valid) : 0u;
	const int c = __popc(mask);
	int incl = c;                                    /* inclusive scan inside the warp, then across the 8 warps */
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int v = __shfl_up_sync(FULL, incl, d);
		if ((int) (threadIdx.x & 31) >= d)
			incl += v;
	}
	__shared__ int wsum[NL_THREADS / 32];
	if ((threadIdx.x & 31) == 31)
		wsum[threadIdx.x >> 5] = incl;
	__syncthreads();
	int before = incl - c;
	for (int k = 0; k < (int) (threadIdx.x >> 5); k++)
		before += wsum[k];
	unsigned slot = t.block_cnt[blockIdx.x] + (unsigned) before;
	while (mask) {
		const int b = __ffs(mask) - 1;
		mask &= mask - 1;
		if (slot < t.nl_cap)
			t.nl[slot] = (uint32_t) (base + b);
		slot++;
	}
}

/* ---- lines --------------------------------------------------------------------------------------------------- */
struct Line {
	const uint8_t *p;
	int len;            /* CR stripped (linebuf.c:82-85) */
	int raw;            /* bytes up to the '\n' */
};
__device__ __forceinline__ Line get_line(const TextView &t, unsigned k) {
	const unsigned s = k ? t.nl[k - 1] + 1u : 0u, e = t.nl[k];
	Line l;
	l.p = t.text + s;
	l.raw = (int) (e - s);
	l.len = l.raw;
	if (l.len > 0 && l.p[l.len - 1] == '\r')
		l.len--;
	return l;
}

__global__ void fq_geometry(TextView tf, TextView tr, ParseState *st, unsigned max_records) {
	/* launched after nl_write with a grid-stride over the records: longest sequence line */
	const unsigned nf = min(st->nl_total[0], tf.nl_cap), nr = min(st->nl_total[1], tr.nl_cap);
	const unsigned records = min(min(nf, nr) / 4, max_records);
	unsigned mf = 0, mr = 0;
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < records; i += gridDim.x * blockDim.x) {
		mf = max(mf, (unsigned) get_line(tf, 4 * i + 1).len);
		mr = max(mr, (unsigned) get_line(tr, 4 * i + 1).len);
	}
	mf = __reduce_max_sync(FULL, mf);
	mr = __reduce_max_sync(FULL, mr);
	if ((threadIdx.x & 31) == 0) {
		if (mf) atomicMax(&st->max_len[0], min(mf, (unsigned) PB_MAX_LEN));
		if (mr) atomicMax(&st->max_len[1], min(mr, (unsigned) PB_MAX_LEN));
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		st->records = records;
		st->consumed[0] = records ? (unsigned long long) tf.nl[4 * records - 1] + 1ull : 0ull;
		st->consumed[1] = records ? (unsigned long long) tr.nl[4 * records - 1] + 1ull : 0ull;
	}
}

__global__ void fq_stride(ParseState *st) {
	const unsigned f = st->max_len[0], r = st->max_len[1];
	const unsigned b = ((f + 7) / 8) * 4 + ((r + 7) / 8) * 4 + ((f + 3) / 4) * 4 + ((r + 3) / 4) * 4;
	st->stride16 = (b + 15) / 16;
}

/* ---- identifiers: seqid.c:136-285, one thread per record -------------------------------------------------------- */
struct Cursor {
	const uint8_t *s;       /* the header as a C string: a NUL (or the end of the line, `len`) ends it */
	int pos, len;
	bool terminated;        /* s[len] == 0: no bounds test needed */
	__device__ __forceinline__ int cur() const { return (terminated || pos < len) ? (int) s[pos] : 0; }
};
__device__ __forceinline__ bool is_delim(int c) {
	return c == 0 || c == ':' || c == '#' || c == '/' || c == ' ';
}
struct IdFields {
	int inst_off, inst_len, run_off, run_len, fc_off, fc_len, tag_off, tag_len;
	int lane, tile, x, y, sra, fmt;
};
__device__ __forceinline__ bool take_str(Cursor &c, int &off, int &len, int cap) {
	if (c.cur() == 0)
		return false;
	off = c.pos;
	while (!is_delim(c.cur()))
		c.pos++;
	len = c.pos - off;
	return len <= cap;       /* a field of 101 characters overruns the reference's struct member (seqid.c:153); refused */
}
__device__ __forceinline__ bool take_int(Cursor &c, int &value) {
	if (c.cur() == 0)
		return false;
	unsigned v = 0;
	while (!is_delim(c.cur())) {
		const int ch = c.cur();
		if (ch < '0' || ch > '9')
			return false;
		v = 10u * v + (unsigned) (ch - '0');
		c.pos++;
	}
	value = (int) v;
	return true;
}
__device__ __forceinline__ bool take_sra_int(Cursor &c, int &value) {
	unsigned v = 0;
	for (int ch = c.cur(); ch != 0 && ch != '.' && ch != ' '; ch = c.cur()) {
		if (ch < '0' || ch > '9')
			return false;
		v = 10u * v + (unsigned) (ch - '0');
		c.pos++;
	}
	value = (int) v;
	return true;
}
__device__ __forceinline__ bool push(Cursor &c) {
	if (c.cur() == 0)
		return false;
	c.pos++;
	return true;
}
__device__ __forceinline__ bool take_tag(Cursor &c, int &off, int &len) {
	off = c.pos;
	while (!is_delim(c.cur()))
		c.pos++;
	len = c.pos - off;
	return len <= PANDA_TAG_LEN;
}
__device__ __forceinline__ bool tag_policy_ok(int tag_len, int policy) {
	if (policy == PB_TAG_OPTIONAL)
		return true;
	return policy == (tag_len == 0 ? PB_TAG_ABSENT : PB_TAG_PRESENT);
}

/* returns the direction (0 = failure).  hdr/len: the header line after its first character, ended by a NUL if it has one */
__device__ int parse_id(const uint8_t *hdr, int len, int policy, IdFields &f, bool terminated) {
	Cursor c = { hdr, 0, len, terminated };
	/* one pass: where the C string ends, whether it holds a '/', and the ':' before the first '#' (seqid.c:170-178) */
	bool slash = false, hashed = false;
	int colons = 0;
	for (int k = 0; k < len; k++) {
		const int ch = hdr[k];
		if (ch == 0) {
			c.len = k;
			break;
		}
		slash |= ch == '/';
		hashed |= ch == '#';
		colons += (ch == ':' && !hashed) ? 1 : 0;
	}
	len = c.len;
	f.inst_off = f.inst_len = f.run_off = f.run_len = f.fc_off = f.fc_len = f.tag_off = f.tag_len = 0;
	f.lane = f.tile = f.x = f.y = f.sra = 0;
	int v;
	if (len > 3 && (hdr[0] == 'E' || hdr[0] == 'S') && hdr[1] == 'R' && hdr[2] == 'R') {
		f.fmt = hdr[0] == 'S' ? PB_IDFMT_SRA : PB_IDFMT_EBI_SRA;
		c.pos = 3;
		if (!take_sra_int(c, v) || !push(c))
			return 0;
		f.sra = v;
		if (!take_sra_int(c, v) || !push(c))
			return 0;
		f.lane = v;
		if (!push(c))
			return 0;
		return 1;
	}
	if (slash) {
		if (colons == 6) {
			f.fmt = PB_IDFMT_CASAVA_CONVERTED;
			if (!take_str(c, f.inst_off, f.inst_len, 100) || !push(c)) return 0;
			if (!take_str(c, f.run_off, f.run_len, 100) || !push(c)) return 0;
			if (!take_str(c, f.fc_off, f.fc_len, 100) || !push(c)) return 0;
		} else {
			f.fmt = PB_IDFMT_CASAVA_1_4;
			if (!take_str(c, f.inst_off, f.inst_len, 100) || !push(c)) return 0;
		}
		if (!take_int(c, f.lane) || !push(c)) return 0;
		if (!take_int(c, f.tile) || !push(c)) return 0;
		if (!take_int(c, f.x) || !push(c)) return 0;
		if (!take_int(c, f.y) || !push(c)) return 0;
		if (hdr[c.pos - 1] == '#') {
			if (!take_tag(c, f.tag_off, f.tag_len) || !push(c))
				return 0;
		}
		if (!tag_policy_ok(f.tag_len, policy))
			return 0;
		if (!take_int(c, v))
			return 0;
		return v;
	}
	f.fmt = PB_IDFMT_CASAVA_1_7;
	int mate, skip_off, skip_len;
	if (!take_str(c, f.inst_off, f.inst_len, 100) || !push(c)) return 0;
	if (!take_str(c, f.run_off, f.run_len, 100) || !push(c)) return 0;
	if (!take_str(c, f.fc_off, f.fc_len, 100) || !push(c)) return 0;
	if (!take_int(c, f.lane) || !push(c)) return 0;
	if (!take_int(c, f.tile) || !push(c)) return 0;
	if (!take_int(c, f.x) || !push(c)) return 0;
	if (!take_int(c, f.y) || !push(c)) return 0;
	if (!take_int(c, mate) || !push(c)) return 0;
	if (!take_str(c, skip_off, skip_len, 1 << 20) || !push(c)) return 0;
	if (!take_int(c, v) || !push(c)) return 0;
	if (!take_tag(c, f.tag_off, f.tag_len))
		return 0;
	if (!tag_policy_ok(f.tag_len, policy))
		return 0;
	return mate;
}

__device__ __forceinline__ bool same_str(const uint8_t *a, int aoff, int alen, const uint8_t *b, int boff, int blen) {
	if (alen != blen)
		return false;
	for (int k = 0; k < alen; k++)
		if (a[aoff + k] != b[boff + k])
			return false;
	return true;
}

__device__ __forceinline__ void report(ParseState *st, unsigned record, int stage, int code) {
	atomicMin(&st->err_key, ((unsigned long long) record << 8) | ((unsigned long long) stage << 4) | (unsigned long long) code);
}

/* One thread parses one record's two header lines.  The lines of a CTA's 128 records are first copied into shared
 * memory by whole warps (coalesced), so the character-at-a-time parser never waits on HBM; a header longer than
 * ID_STAGE characters is parsed in place. */
constexpr int ID_THREADS = 128;
constexpr int ID_STAGE = 124;                      /* characters staged per header (after the first character) */
constexpr int ID_ROW = 132;                        /* row stride in bytes: 33 words, so the threads of a warp hit different banks */

__global__ void __launch_bounds__(ID_THREADS) fq_ids(TextView tf, TextView tr, ParseState *st, int policy, pb_seq_id *ids) {
	__shared__ __align__(16) uint8_t stage[2][ID_THREADS][ID_ROW];
	__shared__ unsigned s_start[2][ID_THREADS];
	__shared__ int s_len[2][ID_THREADS], s_raw[2][ID_THREADS];
	const unsigned records = st->records;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (unsigned base = blockIdx.x * ID_THREADS; base < records; base += gridDim.x * ID_THREADS) {
		const unsigned i = base + threadIdx.x;
		__syncthreads();                           /* the previous round's rows are no longer read */
		if (i < records) {
#pragma unroll
			for (int f = 0; f < 2; f++) {
				const Line h = get_line(f ? tr : tf, 4 * i);
				s_start[f][threadIdx.x] = (unsigned) (h.p - (f ? tr.text : tf.text));
				s_len[f][threadIdx.x] = h.len;
				s_raw[f][threadIdx.x] = h.raw;
			}
		}
		__syncthreads();
		/* warp w copies the headers of records 32w .. 32w+31: one coalesced 132-byte row of aligned words per header, four
		 * records' loads in flight at a time; the parser reads its row at the header's offset inside the first word */
		for (int r0 = 0; r0 < 32; r0 += 4) {
			uint32_t w[4][2], w32[4][2];
#pragma unroll
			for (int u = 0; u < 4; u++)
#pragma unroll
				for (int f = 0; f < 2; f++) {
					const int slot = warp * 32 + r0 + u;
					w[u][f] = w32[u][f] = 0;
					if (base + slot < records && s_len[f][slot] > 0) {
						const TextView &t = f ? tr : tf;
						const unsigned long long first = ((unsigned long long) s_start[f][slot] + 1ull) & ~3ull;
						const uint32_t *src = reinterpret_cast<const uint32_t *>(t.text + first);
						if (first + 4ull * lane + 4ull <= t.bytes)
							w[u][f] = src[lane];
						if (lane == 0 && first + 132ull <= t.bytes)
							w32[u][f] = src[32];
					}
				}
#pragma unroll
			for (int u = 0; u < 4; u++)
#pragma unroll
				for (int f = 0; f < 2; f++) {
					uint32_t *row = reinterpret_cast<uint32_t *>(stage[f][warp * 32 + r0 + u]);
					row[lane] = w[u][f];
					if (lane == 0)
						row[32] = w32[u][f];
				}
		}
		__syncthreads();
		if (i >= records)
			continue;
		IdFields f, r;
		pb_seq_id out;
		memset(&out, 0, sizeof out);
		const int flen = s_len[0][threadIdx.x], rlen = s_len[1][threadIdx.x], fraw = s_raw[0][threadIdx.x], rraw = s_raw[1][threadIdx.x];
		const uint8_t *gf = tf.text + s_start[0][threadIdx.x] + 1, *gr = tr.text + s_start[1][threadIdx.x] + 1;
		const uint8_t *a = flen - 1 <= ID_STAGE ? stage[0][threadIdx.x] + ((s_start[0][threadIdx.x] + 1u) & 3u) : gf;
		const uint8_t *b = rlen - 1 <= ID_STAGE ? stage[1][threadIdx.x] + ((s_start[1][threadIdx.x] + 1u) & 3u) : gr;
		/* fastq.c:125 hands the parser `line + 1` without looking at the first character */
		const bool fterm = flen - 1 <= ID_STAGE, rterm = rlen - 1 <= ID_STAGE;
		if (fterm && flen > 0)
			const_cast<uint8_t *>(a)[flen - 1] = 0;
		if (rterm && rlen > 0)
			const_cast<uint8_t *>(b)[rlen - 1] = 0;
		const int fdir = (fraw < PB_FQ_LINE_MAX && flen > 0) ? parse_id(a, flen - 1, policy, f, fterm) : 0;
		if (fdir == 0) {
			report(st, i, 0, fraw >= PB_FQ_LINE_MAX ? PB_FQ_LINE_TOO_LONG : PB_FQ_ID_PARSE_FAILURE);
		} else {
			const int rdir = (rraw < PB_FQ_LINE_MAX && rlen > 0) ? parse_id(b, rlen - 1, policy, r, rterm) : 0;
			if (rdir == 0) {
				report(st, i, 1, rraw >= PB_FQ_LINE_MAX ? PB_FQ_LINE_TOO_LONG : PB_FQ_ID_PARSE_FAILURE);
			} else {
				/* panda_seqid_equal (seqid.c:95-99); the SRA instrument is "%cRR%d": same letter (fmt) and number */
				bool eq = f.lane == r.lane && f.tile == r.tile && f.x == r.x && f.y == r.y && f.sra == r.sra
					&& ((f.fmt == PB_IDFMT_SRA || f.fmt == PB_IDFMT_EBI_SRA) == (r.fmt == PB_IDFMT_SRA || r.fmt == PB_IDFMT_EBI_SRA))
					&& ((f.fmt != PB_IDFMT_SRA && f.fmt != PB_IDFMT_EBI_SRA) || f.fmt == r.fmt)
					&& same_str(a, f.inst_off, f.inst_len, b, r.inst_off, r.inst_len)
					&& same_str(a, f.run_off, f.run_len, b, r.run_off, r.run_len)
					&& same_str(a, f.fc_off, f.fc_len, b, r.fc_off, r.fc_len)
					&& same_str(a, f.tag_off, min(f.tag_len, PANDA_TAG_LEN), b, r.tag_off, min(r.tag_len, PANDA_TAG_LEN));
				const bool directional = f.fmt != PB_IDFMT_SRA && f.fmt != PB_IDFMT_EBI_SRA;      /* seqid.c:43-46 */
				if (!eq || (directional && rdir == fdir))
					report(st, i, 2, PB_FQ_NOT_PAIRED);
			}
			out.hdr_off = (uint32_t) (gf - tf.text);
			out.hdr_len = (uint16_t) (flen - 1);
			out.fmt = (uint8_t) f.fmt;
			out.inst_off = (uint16_t) f.inst_off; out.inst_len = (uint16_t) f.inst_len;
			out.run_off = (uint16_t) f.run_off; out.run_len = (uint16_t) f.run_len;
			out.fc_off = (uint16_t) f.fc_off; out.fc_len = (uint16_t) f.fc_len;
			out.tag_off = (uint16_t) f.tag_off; out.tag_len = (uint16_t) f.tag_len;
			out.lane = f.lane; out.tile = f.tile; out.x = f.x; out.y = f.y; out.sra = f.sra; out.mate = fdir;
		}
		if (ids) {
			const uint4 *src = reinterpret_cast<const uint4 *>(&out);
			uint4 *dst = reinterpret_cast<uint4 *>(&ids[i]);
			dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
		}
	}
}

/* ---- reads: fastq.c:44-102, one warp per record --------------------------------------------------------------- */
/* nt.c:48-118: letter & 0x1F -> 4-bit code, two codes per byte of a 64-bit constant pair */
__device__ __forceinline__ unsigned letter_code(unsigned ch, bool complement) {
	/* index (ch & 31):  0 1  2 3  4 5 6 7  8 9 10 11 12 13 14 15 | 16 17 18 19 20 21 22 23 24 25 26..31 */
	/* iupac_forward:    0 1 14 2 13 0 0 4 11 0  0 12  0  3 15  0 |  0  0  5  6  8  8  7  9 15 10  0      */
	const unsigned idx = ch & 31u;
	const unsigned long long tab = idx < 16u ? 0x0F30C00B400D2E10ull : 0x000000AF97886500ull;   /* nibble k = code of index k */
	unsigned v = (unsigned) ((tab >> (4u * (idx & 15u))) & 15ull);
	if (complement)
		v = __brev(v) >> 28;                /* A<->T, C<->G: the four bits reversed (nt.c:27-44, 85-118) */
	return v;
}

__device__ __forceinline__ int to_index(int ch, int qualmin) {      /* fastq.c:44, on a signed char */
	if (ch < qualmin)
		return 0;
	return (ch > qualmin + PB_PHREDMAX ? PB_PHREDMAX : ch) - qualmin;
}

/* One record with every byte of its two reads held in registers: NI characters per lane and line.  The line index,
 * the CR / '+' probes and the characters are each fetched in ONE batch of independent loads. */
template <int NI>
__device__ __forceinline__ void reads_record(const TextView &tf, const TextView &tr, ParseState *st, int qualmin, unsigned i, unsigned stride16,
                                             uint8_t *reads, unsigned long long reads_cap, pb_pair_meta *meta, int lane,
                                             const uint8_t *__restrict__ lut) {
	/* lanes 0..4: forward newline positions 4i-1 .. 4i+3; lanes 8..12: the reverse ones */
	const bool rev_lane = lane >= 8;
	const int kk = (lane & 7);
	unsigned pos = 0;
	if (kk < 5 && lane < 13) {
		const TextView &t = rev_lane ? tr : tf;
		const long long q = (long long) 4 * i - 1 + kk;
		pos = q < 0 ? 0xFFFFFFFFu : t.nl[q];
	}
	/* line k of the record (1 = sequence, 2 = '+', 3 = quality) of file f: start = pos[f][k-1]+1, end = pos[f][k] */
	unsigned ls[2][3], le[2][3];
#pragma unroll
	for (int f = 0; f < 2; f++)
#pragma unroll
		for (int k = 0; k < 3; k++) {
			ls[f][k] = __shfl_sync(FULL, pos, 8 * f + k + 1) + 1u;
			le[f][k] = __shfl_sync(FULL, pos, 8 * f + k + 2);
		}
	/* one probe per line: its last byte (CR?) and, for the '+' line, its first byte */
	unsigned last = 0, first = 0;
	if (lane < 6) {
		const int f = lane / 3, k = lane % 3;
		const uint8_t *text = f ? tr.text : tf.text;
		unsigned s0 = 0, e0 = 0;
#pragma unroll
		for (int ff = 0; ff < 2; ff++)
#pragma unroll
			for (int k2 = 0; k2 < 3; k2++)
				if (ff == f && k2 == k) {
					s0 = ls[ff][k2];
					e0 = le[ff][k2];
				}
		if (e0 > s0) {
			last = text[e0 - 1];
			first = text[s0];
		}
	}
	int raw[2][3], len[2][3];
#pragma unroll
	for (int f = 0; f < 2; f++)
#pragma unroll
		for (int k = 0; k < 3; k++) {
			raw[f][k] = (int) (le[f][k] - ls[f][k]);
			const unsigned lb = __shfl_sync(FULL, last, 3 * f + k);
			len[f][k] = raw[f][k] - ((raw[f][k] > 0 && lb == '\r') ? 1 : 0);
		}
	const unsigned plus_f = __shfl_sync(FULL, first, 1), plus_r = __shfl_sync(FULL, first, 4);
	const int F = min(len[0][0], PB_MAX_LEN), R = min(len[1][0], PB_MAX_LEN);
	/* every character this lane is responsible for: element j = j0 + 32 n of the packed read (template order for the
	 * reverse read: element j is read position R-1-j) */
	unsigned cs[2][NI], cq[2][NI];
	const uint8_t *fs = tf.text + ls[0][0], *fq = tf.text + ls[0][2], *rs = tr.text + ls[1][0], *rq = tr.text + ls[1][2];
	const bool qual_ok_f = len[0][2] == F, qual_ok_r = len[1][2] == R;
#pragma unroll
	for (int n = 0; n < NI; n++) {
		const int j = lane + 32 * n;
		cs[0][n] = j < F ? fs[j] : 'A';
		cq[0][n] = (j < F && qual_ok_f) ? fq[j] : 1;
		cs[1][n] = j < R ? rs[R - 1 - j] : 'A';
		cq[1][n] = (j < R && qual_ok_r) ? rq[R - 1 - j] : 1;
	}
	/* letters -> 4-bit codes, once (lut[0..31] = iupac_forward, lut[32..63] = iupac_reverse, nt.c:48-118) */
#pragma unroll
	for (int n = 0; n < NI; n++) {
		cs[0][n] = lut[cs[0][n] & 31u];
		cs[1][n] = lut[32u + (cs[1][n] & 31u)];
	}
	/* checks, in the reference's order (fastq.c:57-99), forward read first */
	int code = PB_FQ_OK, stage = 3;
#pragma unroll
	for (int f = 0; f < 2 && code == PB_FQ_OK; f++) {
		stage = 3 + f;
		bool bad = false, nul = false;
#pragma unroll
		for (int n = 0; n < NI; n++) {
			bad |= cs[f][n] == 0u;
			nul |= cq[f][n] == 0u;
		}
		const unsigned plus = f ? plus_r : plus_f;
		if (raw[f][0] >= PB_FQ_LINE_MAX)
			code = PB_FQ_LINE_TOO_LONG;
		else if (__any_sync(FULL, bad))
			code = PB_FQ_BAD_NT;
		else if (raw[f][1] >= PB_FQ_LINE_MAX)
			code = PB_FQ_LINE_TOO_LONG;
		else if (len[f][1] == 0 || plus != '+')
			code = letter_code(len[f][1] == 0 ? 0u : plus, f == 1) != 0u ? PB_FQ_READ_TOO_LONG : PB_FQ_PARSE_FAILURE;
		else if (raw[f][2] >= PB_FQ_LINE_MAX)
			code = PB_FQ_LINE_TOO_LONG;
		else if (len[f][2] != (f ? R : F) || __any_sync(FULL, nul))
			code = PB_FQ_NO_QUALITY_INFO;
	}
	pb_pair_meta m;
	m.off16 = i * stride16;
	m.flen = 0xFFFF;
	m.rlen = 0;
	const unsigned long long rec_at = (unsigned long long) i * stride16 * 16ull;
	if (code != PB_FQ_OK) {
		if (lane == 0)
			report(st, i, stage, code);
	} else if (F > 0 && rec_at + (unsigned long long) stride16 * 16ull <= reads_cap) {     /* fastq.c:176: an empty forward read is dropped */
		uint8_t *rec = reads + rec_at;
		const int fwb = ((F + 7) / 8) * 4, rwb = ((R + 7) / 8) * 4, fqb = ((F + 3) / 4) * 4, rqb = ((R + 3) / 4) * 4;
		uint8_t *nt_out[2] = { rec, rec + fwb }, *q_out[2] = { rec + fwb + rwb, rec + fwb + rwb + fqb };
		const int ntb[2] = { fwb, rwb }, qb[2] = { fqb, rqb }, ln[2] = { F, R };
#pragma unroll
		for (int f = 0; f < 2; f++)
#pragma unroll
			for (int n = 0; n < NI; n++) {
				const int j = lane + 32 * n;
				const unsigned c = j < ln[f] ? cs[f][n] : 0u;
				const unsigned hi = __shfl_down_sync(FULL, c, 1);
				if ((lane & 1) == 0 && j < 2 * ntb[f])
					nt_out[f][j >> 1] = (uint8_t) (c | (hi << 4));
				if (j < qb[f])
					q_out[f][j] = (uint8_t) (j < ln[f] ? to_index((int) (signed char) cq[f][n], qualmin) : 0);
			}
		const int used = fwb + rwb + fqb + rqb, total = (used + 15) & ~15;
		for (int k = used + lane; k < total; k += 32)
			rec[k] = 0;
		m.flen = (uint16_t) F;
		m.rlen = (uint16_t) R;
	}
	if (lane == 0)
		meta[i] = m;
}

/* NI = characters per lane and line; the host launches the three instantiations back to back and the two whose
 * length class does not match the chunk return at once (the longest read is known only on the device). */
template <int NI>
__global__ void __launch_bounds__(128, (NI <= 5 ? 8 : (NI <= 10 ? 5 : 3))) fq_reads(TextView tf, TextView tr, ParseState *st, int qualmin,
                                                                                   uint8_t *reads, unsigned long long reads_cap, pb_pair_meta *meta) {
	const unsigned longest = max(st->max_len[0], st->max_len[1]);
	const int want = longest <= 160 ? 5 : (longest <= 320 ? 10 : 15);
	if (want != NI)
		return;
	__shared__ uint8_t lut[64];
	if (threadIdx.x < 64)
		lut[threadIdx.x] = (uint8_t) letter_code(threadIdx.x & 31u, threadIdx.x >= 32);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const unsigned records = st->records;
	const unsigned stride16 = st->stride16;
	const unsigned warps = (gridDim.x * blockDim.x) >> 5;
	for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < records; i += warps)
		reads_record<NI>(tf, tr, st, qualmin, i, stride16, reads, reads_cap, meta, lane, lut);
}

__global__ void fq_finish(ParseState *st, const pb_pair_meta *meta) {
	/* single CTA: the record that ended the stream, and the pairs delivered before it */
	__shared__ unsigned long long s_limit;
	__shared__ unsigned s_skipped;
	if (threadIdx.x == 0) {
		const unsigned long long key = st->err_key;
		s_limit = key == NO_ERROR_KEY ? (unsigned long long) st->records : (key >> 8);
		st->error = key == NO_ERROR_KEY ? PB_FQ_OK : (int) (key & 15ull);
		s_skipped = 0;
	}
	__syncthreads();
	unsigned skipped = 0;
	for (unsigned long long i = threadIdx.x; i < s_limit; i += blockDim.x)
		skipped += meta[i].flen == 0xFFFF;
	skipped = __reduce_add_sync(FULL, skipped);
	if ((threadIdx.x & 31) == 0 && skipped)
		atomicAdd(&s_skipped, skipped);
	__syncthreads();
	if (threadIdx.x == 0) {
		st->limit = s_limit;
		st->pairs = s_limit - s_skipped;
	}
}

/* ---- output: output.c:85-126 -------------------------------------------------------------------------------- */
__device__ __forceinline__ int int_chars(int v) {          /* strlen of "%d" */
	const unsigned u = v < 0 ? 0u - (unsigned) v : (unsigned) v;
	return (v < 0 ? 2 : 1) + (u >= 10u) + (u >= 100u) + (u >= 1000u) + (u >= 10000u) + (u >= 100000u) + (u >= 1000000u)
		+ (u >= 10000000u) + (u >= 100000000u) + (u >= 1000000000u);
}
__device__ __forceinline__ int put_int(char *dst, int v, int n) {      /* n = int_chars(v) */
	unsigned u = v < 0 ? 0u - (unsigned) v : (unsigned) v;
	for (int k = n - 1; k >= (v < 0 ? 1 : 0); k--) {
		dst[k] = (char) ('0' + u % 10u);
		u /= 10u;
	}
	if (v < 0)
		dst[0] = '-';
	return n;
}
/* round(v * 10^6) to nearest, ties to even, on the exact binary value -- what printf("%f") prints; v >= 0 finite */
__device__ __forceinline__ unsigned long long scaled_micro(double v) {
	const unsigned long long bits = (unsigned long long) __double_as_longlong(v);
	int e = (int) ((bits >> 52) & 0x7FFull);
	unsigned long long m = bits & ((1ull << 52) - 1ull);
	if (e == 0)
		e = 1;
	else
		m |= 1ull << 52;
	const int sh = 1075 - e;                     /* v = m / 2^sh */
	unsigned __int128 N = (unsigned __int128) m * 1000000u;
	if (sh <= 0)
		return (unsigned long long) (N << (-sh));   /* v < 2^63 / 10^6 assumed: exp(quality) <= 1 */
	if (sh > 100)
		return 0ull;
	const unsigned __int128 q = N >> sh, rem = N & ((((unsigned __int128) 1) << sh) - 1), half = ((unsigned __int128) 1) << (sh - 1);
	unsigned long long out = (unsigned long long) q;
	if (rem > half || (rem == half && (out & 1ull)))
		out++;
	return out;
}
__device__ __forceinline__ int f6_chars(unsigned long long micro) {
	if (micro < 10000000ull)                    /* exp(quality) <= 1: always this branch in practice */
		return 8;                               /* "d.dddddd" */
	unsigned long long ip = micro / 1000000ull;
	int n = 8;
	while (ip >= 10ull) {
		ip /= 10ull;
		n++;
	}
	return n;
}
__device__ __forceinline__ int put_f6(char *dst, unsigned long long micro) {
	if (micro < 10000000ull) {                  /* 32-bit arithmetic: 64-bit division is a subroutine on the GPU */
		unsigned v = (unsigned) micro;
#pragma unroll
		for (int k = 7; k > 1; k--) {
			dst[k] = (char) ('0' + v % 10u);
			v /= 10u;
		}
		dst[1] = '.';
		dst[0] = (char) ('0' + v);
		return 8;
	}
	const int n = f6_chars(micro);
	unsigned long long ip = micro / 1000000ull;
	unsigned fr = (unsigned) (micro % 1000000ull);
	for (int k = n - 1; k > n - 7; k--) {
		dst[k] = (char) ('0' + fr % 10u);
		fr /= 10u;
	}
	dst[n - 7] = '.';
	for (int k = n - 8; k >= 0; k--) {
		dst[k] = (char) ('0' + ip % 10ull);
		ip /= 10ull;
	}
	return n;
}
__device__ __forceinline__ int sra_chars(int v) { return 3 + int_chars(v); }

/* nt.c:126-150 */
__device__ __forceinline__ int result_phred(double p, const double *score) {
	int lower = 0, upper = PB_PHREDMAX;
	if (p <= score[0])
		return 1;
	while (lower < upper) {
		const int mid = lower + (upper - lower) / 2;
		const double s = score[mid];
		if (s == p)
			return mid;
		if (mid == lower)
			return lower;
		if (s > p)
			upper = mid;
		else if (s < p)
			lower = mid + 1;
	}
	return lower;
}

__device__ __forceinline__ bool emitted(const pb_pair_result &r) {
	return r.status == PB_PAIR_OK && r.seq_len > 0;        /* output.c:88,108 */
}

/* header text length: "%s:%s:%s:%d:%d:%d:%d:%s" (seqid.c:121-128) + ";%f" */
__device__ __forceinline__ int header_chars(const pb_seq_id &id, unsigned long long micro) {
	const bool sra = id.fmt == PB_IDFMT_SRA || id.fmt == PB_IDFMT_EBI_SRA;
	return 1 + (sra ? sra_chars(id.sra) : (int) id.inst_len) + 1 + id.run_len + 1 + id.fc_len + 1 + int_chars(id.lane) + 1 + int_chars(id.tile)
		+ 1 + int_chars(id.x) + 1 + int_chars(id.y) + 1 + min((int) id.tag_len, PANDA_TAG_LEN) + 1 + f6_chars(micro) + 1;
}

__global__ void fmt_length(int n, int fastq, const pb_pair_result *__restrict__ res, const pb_seq_id *__restrict__ ids,
                           uint32_t *__restrict__ len_out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const pb_pair_result r = res[i];
	unsigned len = 0;
	if (emitted(r)) {
		const pb_seq_id id = ids[i];
		len = (unsigned) header_chars(id, scaled_micro(exp(r.quality))) + r.seq_len + 1u;
		if (fastq)
			len += 2u + r.seq_len + 1u;
	}
	len_out[i] = len;
}

/* generic exclusive scan of n u32 (in place) with a 64-bit total: per-tile sums, scan of the sums, apply */
constexpr int SCAN_TILE = 2048;
__global__ void __launch_bounds__(256) scan_tile_sums(int n, const uint32_t *__restrict__ v, unsigned long long *__restrict__ tile_sum) {
	const int base = blockIdx.x * SCAN_TILE;
	unsigned long long s = 0;
	for (int k = threadIdx.x; k < SCAN_TILE && base + k < n; k += 256)
		s += v[base + k];
	__shared__ unsigned long long sh[256];
	sh[threadIdx.x] = s;
	__syncthreads();
	for (int d = 128; d > 0; d >>= 1) {
		if ((int) threadIdx.x < d)
			sh[threadIdx.x] += sh[threadIdx.x + d];
		__syncthreads();
	}
	if (threadIdx.x == 0)
		tile_sum[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(1024) scan_tiles(int ntiles, unsigned long long *tile_sum, unsigned long long *total) {
	__shared__ unsigned long long part[1024];
	const int per = (ntiles + 1023) / 1024;
	const int lo = min((int) threadIdx.x * per, ntiles), hi = min(lo + per, ntiles);
	unsigned long long s = 0;
	for (int k = lo; k < hi; k++)
		s += tile_sum[k];
	part[threadIdx.x] = s;
	__syncthreads();
	for (int d = 1; d < 1024; d <<= 1) {
		unsigned long long v = (int) threadIdx.x >= d ? part[threadIdx.x - d] : 0ull;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	unsigned long long run = part[threadIdx.x] - s;
	for (int k = lo; k < hi; k++) {
		const unsigned long long c = tile_sum[k];
		tile_sum[k] = run;
		run += c;
	}
	if (threadIdx.x == 1023)
		*total = part[1023];
}

/* off[i] = exclusive prefix of v[0..i) */
__global__ void __launch_bounds__(256) scan_apply(int n, const uint32_t *__restrict__ v, const unsigned long long *__restrict__ tile_off,
                                                  unsigned long long *__restrict__ off) {
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 8;
	unsigned e[8];
	unsigned long long s = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) {
		e[k] = base + k < n ? v[base + k] : 0u;
		s += e[k];
	}
	unsigned long long incl = s;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long t = __shfl_up_sync(FULL, incl, d);
		if ((int) (threadIdx.x & 31) >= d)
			incl += t;
	}
	__shared__ unsigned long long wsum[8];
	if ((threadIdx.x & 31) == 31)
		wsum[threadIdx.x >> 5] = incl;
	__syncthreads();
	unsigned long long run = tile_off[blockIdx.x] + incl - s;
	for (int k = 0; k < (int) (threadIdx.x >> 5); k++)
		run += wsum[k];
#pragma unroll
	for (int k = 0; k < 8; k++) {
		if (base + k < n)
			off[base + k] = run;
		run += e[k];
	}
}

/* one warp per record */
__global__ void __launch_bounds__(256) fmt_write(int n, int fastq, const pb_pair_result *__restrict__ res,
                                                 const uint8_t *__restrict__ seq_nt, const double *__restrict__ seq_p, long long seq_stride,
                                                 const pb_seq_id *__restrict__ ids, const uint8_t *__restrict__ fwd_text,
                                                 const uint32_t *__restrict__ len_in, const unsigned long long *__restrict__ rec_off,
                                                 const double *__restrict__ score, char *__restrict__ text, unsigned long long capacity) {
	__shared__ double s_score[PB_NQ];
	for (int k = threadIdx.x; k < PB_NQ; k += blockDim.x)
		s_score[k] = score[k];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (i >= n)
		return;
	const unsigned mylen = len_in[i];
	if (mylen == 0)
		return;
	const unsigned long long off = rec_off[i];
	if (off + mylen > capacity)
		return;
	const pb_pair_result r = res[i];
	const pb_seq_id id = ids[i];
	const uint8_t *src = fwd_text + id.hdr_off;
	const bool sra = id.fmt == PB_IDFMT_SRA || id.fmt == PB_IDFMT_EBI_SRA;
	/* where every piece of "%s:%s:%s:%d:%d:%d:%d:%s;%f\n" goes (all lanes compute the same layout) */
	const unsigned long long micro = scaled_micro(exp(r.quality));
	const int l_inst = sra ? sra_chars(id.sra) : (int) id.inst_len, l_tag = min((int) id.tag_len, PANDA_TAG_LEN);
	const int p_inst = 1, p_run = p_inst + l_inst + 1, p_fc = p_run + id.run_len + 1, p_lane = p_fc + id.fc_len + 1;
	const int l_lane = int_chars(id.lane), l_tile = int_chars(id.tile), l_x = int_chars(id.x), l_y = int_chars(id.y);
	const int p_tile = p_lane + l_lane + 1, p_x = p_tile + l_tile + 1, p_y = p_x + l_x + 1;
	const int p_tag = p_y + l_y + 1, p_q = p_tag + l_tag + 1, hl = p_q + f6_chars(micro) + 1;
	char *dst = text + off;
	if (lane < 10) {                    /* the punctuation, one character per lane */
		const int at[10] = { 0, p_run - 1, p_fc - 1, p_lane - 1, p_tile - 1, p_x - 1, p_y - 1, p_tag - 1, p_q - 1, hl - 1 };
		int where = 0;
#pragma unroll
		for (int k = 0; k < 10; k++)
			if (lane == k)
				where = at[k];
		dst[where] = lane == 0 ? (fastq ? '@' : '>') : (lane == 8 ? ';' : (lane == 9 ? '\n' : ':'));
	} else if (lane < 14) {             /* the four integers */
		const int v = lane == 10 ? id.lane : (lane == 11 ? id.tile : (lane == 12 ? id.x : id.y));
		const int at = lane == 10 ? p_lane : (lane == 11 ? p_tile : (lane == 12 ? p_x : p_y));
		put_int(dst + at, v, lane == 10 ? l_lane : (lane == 11 ? l_tile : (lane == 12 ? l_x : l_y)));
	} else if (lane == 14) {
		put_f6(dst + p_q, micro);
	} else if (lane == 15 && sra) {
		dst[p_inst] = id.fmt == PB_IDFMT_SRA ? 'S' : 'E';
		dst[p_inst + 1] = 'R';
		dst[p_inst + 2] = 'R';
		put_int(dst + p_inst + 3, id.sra, l_inst - 3);
	}
	if (!sra && id.run_off == id.inst_off + id.inst_len + 1 && id.fc_off == id.run_off + id.run_len + 1) {
		/* "instrument:run:flowcell" stand next to each other in the header: one copy, the two separators are rewritten as ':' below */
		const int n = id.inst_len + 1 + id.run_len + 1 + id.fc_len;
		for (int k = lane; k < n; k += 32) {
			const bool sep = k == id.inst_len || k == id.inst_len + 1 + id.run_len;
			dst[p_inst + k] = sep ? ':' : (char) src[id.inst_off + k];
		}
	} else {
		if (!sra)
			for (int k = lane; k < id.inst_len; k += 32)
				dst[p_inst + k] = (char) src[id.inst_off + k];
		for (int k = lane; k < id.run_len; k += 32)
			dst[p_run + k] = (char) src[id.run_off + k];
		for (int k = lane; k < id.fc_len; k += 32)
			dst[p_fc + k] = (char) src[id.fc_off + k];
	}
	for (int k = lane; k < l_tag; k += 32)
		dst[p_tag + k] = (char) src[id.tag_off + k];
	dst += hl;
	const uint8_t *nt = seq_nt + (size_t) i * (size_t) (seq_stride / 2);
	const int L = r.seq_len;
	{
		/* nt.c:25 "NACMGRSVTWYHKDBN": a byte-permute picks the letter of a 3-bit index out of a register pair */
		const uint32_t *nt32 = reinterpret_cast<const uint32_t *>(nt);
		const int nw = (L + 7) >> 3;
		for (int w = lane; w < nw; w += 32) {
			unsigned x = nt32[w];
			const int n = min(L - 8 * w, 8);
#pragma unroll
			for (int t = 0; t < 8; t++) {
				const unsigned c = x & 15u;
				x >>= 4;
				const unsigned lo = __byte_perm(0x4D43414Eu, 0x56535247u, c & 7u), hi = __byte_perm(0x48595754u, 0x4E42444Bu, c & 7u);
				if (t < n)
					dst[8 * w + t] = (char) ((c & 8u) ? hi : lo);
			}
		}
	}
	if (lane == 0)
		dst[L] = '\n';
	if (fastq) {
		dst += L + 1;
		if (lane == 0) {
			dst[0] = '+';
			dst[1] = '\n';
		}
		dst += 2;
		const double *p = seq_p + (size_t) i * (size_t) seq_stride;
#pragma unroll 4
		for (int k = lane; k < L; k += 32)
			dst[k] = (char) (33 + result_phred(p[k], s_score));
		if (lane == 0)
			dst[L] = '\n';
	}
}

}  // namespace pbio
