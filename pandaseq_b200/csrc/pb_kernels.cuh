/* pb_kernels.cuh -- sm_100a kernels for the PANDAseq pair-assembly hot path.
 *
 * assemble_kernel<ML>: one warp owns one read pair at a time (persistent grid,
 * warps stride over the batch).  Per pair, following the reference's align()
 * (assembler.c:48-250) and assemble_seq() (assembler.c:252-348):
 *
 *   stage   the packed record (4-bit nt + 8-bit PHRED, both reads) is pulled from
 *           HBM into this warp's shared-memory stage by ONE bulk async copy
 *           (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) completing on an
 *           mbarrier; the copy for the warp's next pair is in flight while the
 *           current pair is processed (2 stages).
 *   primers panda_compute_offset_qual (offset.c:47-112), when primers are set.
 *   planes  three bit-planes per read (hi/lo bit of the 2-bit k-mer code, is-N)
 *           built with warp ballots: 32 bases -> one 32-bit word per plane.
 *   seed    K1-K3 of align(): forward 8-mers go into a per-warp shared-memory hash
 *           (open addressing, write-then-verify instead of atomics); reverse
 *           8-mers probe it and keep the two lowest forward positions per code
 *           (the observable behaviour of the reference's 65536x2 table, SURVEY.md
 *           §8a "table-free statement"); hits set byte flags per candidate overlap.
 *   score   K4/K5: every flagged overlap (or all, if none was flagged) is scored by
 *           the algorithm's overlap_probability, whole warp per candidate.
 *   merge   K6: the merged read, per-base log p (2x48x48 LUT in shared memory),
 *           quality = sum / len, mismatch / degenerate counts; result record out.
 *
 * Floating point: compiled with --fmad=false.  simple_bayes / flash scores are
 * closed forms of integer counts and are bit-identical to the reference.  The
 * pear / rdp_mle score and the quality sum are sums of LUT entries; the reference
 * adds them left to right, the warp adds 32 partial sums with a shuffle tree:
 * |difference| <= ~3e-13 (measured), tolerance 1e-6 (BASELINE.json north_star).
 */
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "pb_internal.h"

namespace pb {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int NSTAGE = 2;

/* ---- per-warp shared memory layout, sized by the read-length class ML -------- */
template <int ML> struct WarpSmem {
	static constexpr int NTW = (ML + 7) / 8;                 /* nibble words per read */
	static constexpr int STAGE_BYTES = ((2 * NTW * 4 + 2 * ((ML + 3) / 4) * 4) + 15) & ~15;
	static constexpr int SLOTS = (ML <= 160) ? 512 : 1024;   /* >= 2x the most k-mers a read can have */
	static constexpr int PLANE_WORDS = ML / 32 + 2;
	static constexpr int NFLAG = ((2 * ML + 15) & ~15) + 16;
	alignas(128) uint8_t stage[NSTAGE][STAGE_BYTES];
	alignas(16) uint32_t htab[SLOTS];
	alignas(16) uint8_t cflag[NFLAG];
	uint32_t plane[6][PLANE_WORDS];   /* f.hi f.lo f.N r.hi r.lo r.N */
	alignas(8) uint64_t bar[NSTAGE];
	pb_pair_meta meta[NSTAGE];
};

/* ---- mbarrier + bulk copy (PTX; SASS shows SYNCS / UBLKCP) --------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ---- small helpers ---------------------------------------------------------------- */
__device__ __forceinline__ int clampq(int q) {      /* prob.h:23 on a signed char */
	return min(max(q, 0), PB_PHREDMAX);
}
__device__ __forceinline__ unsigned nib(const uint8_t *nt, int i) {
	return (nt[i >> 1] >> ((i & 1) * 4)) & 15u;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int s = 16; s > 0; s >>= 1)
		v += __shfl_xor_sync(FULL, v, s);
	return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
	return __reduce_add_sync(FULL, v);
}
/* bits [pos, pos+32) of a bit array stored as 32-bit words */
__device__ __forceinline__ unsigned window(const uint32_t *w, int pos) {
	int k = pos >> 5;
	return __funnelshift_r(w[k], w[k + 1], pos & 31);
}
__device__ __forceinline__ unsigned hash16(unsigned code, unsigned mask) {
	return ((code * 40503u) >> 4) & mask;
}

/* same arithmetic as pb_record_bytes() in the public header */
__device__ __forceinline__ unsigned record_bytes(unsigned flen, unsigned rlen) {
	unsigned b = ((flen + 7) / 8) * 4 + ((rlen + 7) / 8) * 4 + ((flen + 3) / 4) * 4 + ((rlen + 3) / 4) * 4;
	return (b + 15u) & ~15u;
}

struct PairView {
	const uint8_t *fnt, *rnt;      /* packed nibbles; rnt in template order */
	const int8_t *fq, *rq;         /* raw PHRED chars; rq in template order */
	int F, R;
};

/* offset.c:47-112 for one read, whole warp.  Lane s owns the circular-buffer slot s, s+32, ...
 * (P <= 449 slots, up to 15 per lane, kept in registers is too much -> shared would be needed;
 * instead each START offset is owned by a lane: the alignment that begins at read position s
 * accumulates primer[x] vs read[s+x] for x = 0..P-1 in that order, exactly the order the
 * reference's circular buffer receives its addends, so the sum is bit-identical.)
 *
 * template_order: the read is stored reversed (reverse read), so read position i is element len-1-i.
 * Returns bestindex as the reference does (0 = not found, else 1 + bases consumed).
 * With penalty == 0 the comparison exp(a) > exp(b) is done as a > b (exp is monotone; SURVEY.md
 * §8a a18 measured 0 differences on 600 k reads); with a penalty, CUDA's exp() is used. */
__device__ int primer_offset(const uint8_t *nt, const int8_t *q, int len, bool template_order,
                             const uint8_t *primer, int P, double threshold, double penalty,
                             const double *score, const double *score_err, int lane) {
	if (P > len)
		return 0;
	/* The reference tests slot (index % P) at every index before resetting it; the value it sees at
	 * index >= P is the complete sum for start s = index - P.  Starts s in [0, len-P-1] get tested
	 * (the alignment ending exactly at the read end, s = len-P, is never tested).  For index < P the
	 * slot holds -inf: exp(-inf) - index*penalty can only win when penalty < 0, which the setter forbids. */
	double best = (double) P * threshold;            /* log of bestpr = exp(P*threshold), offset.c:59 */
	if (penalty != 0.0)
		best = exp(best);
	int best_index = 0;
	const int nstart = len - P;                    /* starts 0 .. nstart-1 */
	for (int base = 0; base < nstart; base += 32) {
		int s = base + lane;
		double sum = 0.0;
		bool live = s < nstart;
		if (live) {
			for (int x = 0; x < P; x++) {
				unsigned pn = primer[x];
				if (pn == 15u)
					continue;
				int pos = s + x;
				int el = template_order ? (len - 1 - pos) : pos;
				unsigned b = nib(nt, el);
				int ph = clampq(q[el]);
				sum += (b & pn) ? score[ph] : score_err[ph];
			}
		}
		/* The reference scans starts in increasing order and keeps the first strictly better one:
		 * within a batch that is the maximum value with the lowest start on ties. */
		int index = s + P;                     /* the index at which this slot is examined */
		double val = sum / (double) (index + 1);
		if (penalty != 0.0)
			val = exp(val) - (double) index * penalty;
		if (!live)
			val = -CUDART_INF;
		int who = s;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			double ov = __shfl_xor_sync(FULL, val, d);
			int ow = __shfl_xor_sync(FULL, who, d);
			if (ov > val || (ov == val && ow < who)) {
				val = ov;
				who = ow;
			}
		}
		if (val > best) {
			best = val;
			best_index = who + P + 1;
		}
	}
	return best_index;
}

template <int ML>
__device__ void process_pair(WarpSmem<ML> &ws, const uint8_t *rec, int F, int R,
                             const pb_device_params *__restrict__ prm,
                             const double *__restrict__ s_recon, const double *__restrict__ s_over,
                             const double *__restrict__ s_score, const double *__restrict__ s_score_err,
                             pb_pair_result &res, uint8_t *out_nt, double *out_p, int out_cap, int lane) {
	using WS = WarpSmem<ML>;
	PairView v;
	v.F = F;
	v.R = R;
	const int fw = ((F + 7) / 8) * 4, rw = ((R + 7) / 8) * 4;
	v.fnt = rec;
	v.rnt = rec + fw;
	v.fq = (const int8_t *) (rec + fw + rw);
	v.rq = v.fq + ((F + 3) / 4) * 4;

	res.status = PB_PAIR_OK;
	res.slow = 0;
	res.overlap = res.seq_len = res.mismatches = res.degenerates = res.examined = 0;
	res.fwd_offset = res.rev_offset = 0;
	res.quality = 0.0;
	res.est_prob = 0.0;

	/* assembler.c:255-258 */
	if (F < 2 || R < 2) {
		res.status = PB_PAIR_BADR;
		return;
	}
	const int mo = prm->minoverlap;
	int fo, ro;
	/* assembler.c:262-284 (primers before assembly) */
	if (prm->forward_primer_length > 0) {
		fo = primer_offset(v.fnt, v.fq, F, false, prm->forward_primer, prm->forward_primer_length,
		                   prm->threshold, prm->primer_penalty, s_score, s_score_err, lane);
		if (fo == 0) {
			res.status = PB_PAIR_NOFP;
			return;
		}
		fo--;
	} else {
		fo = prm->forward_trim;
	}
	res.fwd_offset = (uint16_t) fo;
	if (prm->reverse_primer_length > 0) {
		ro = primer_offset(v.rnt, v.rq, R, true, prm->reverse_primer, prm->reverse_primer_length,
		                   prm->threshold, prm->primer_penalty, s_score, s_score_err, lane);
		if (ro == 0) {
			res.status = PB_PAIR_NORP;
			return;
		}
		ro--;
	} else {
		ro = prm->reverse_trim;
	}
	res.rev_offset = (uint16_t) ro;
	/* assembler.c:289-292 */
	if (min(F, R) < mo) {
		res.status = PB_PAIR_BADR;
		return;
	}
	/* align(): assembler.c:73-76 */
	if (mo + fo >= F || mo + ro >= R) {
		res.status = PB_PAIR_NOALGN;
		return;
	}
	/* assembler.c:59,78-84 */
	int maxov;
	if (prm->maxoverlap == 0)
		maxov = min(F, R);
	else
		maxov = min(F + R - mo - fo - ro - 1, prm->maxoverlap);
	const int nbits = (mo <= maxov) ? (maxov - mo + 1) : 1;

	/* ---- bit planes (K1/K2 input): misc.h:41 code T=3 G=2 C=1 else 0; N resets the window ---- */
	const int fwords = (F + 31) >> 5, rwords = (R + 31) >> 5;
	bool anyN = false;
	for (int w = 0; w < fwords; w++) {
		int b = w * 32 + lane;
		unsigned n = (b < F) ? nib(v.fnt, b) : 0u;
		unsigned hi = __ballot_sync(FULL, n == 8u || n == 4u);
		unsigned lo = __ballot_sync(FULL, n == 8u || n == 2u);
		unsigned nn = __ballot_sync(FULL, n == 15u);
		anyN |= nn != 0;
		if (lane == 0) {
			ws.plane[0][w] = hi;
			ws.plane[1][w] = lo;
			ws.plane[2][w] = nn;
		}
	}
	for (int w = 0; w < rwords; w++) {
		int b = w * 32 + lane;
		unsigned n = (b < R) ? nib(v.rnt, b) : 0u;
		unsigned hi = __ballot_sync(FULL, n == 8u || n == 4u);
		unsigned lo = __ballot_sync(FULL, n == 8u || n == 2u);
		unsigned nn = __ballot_sync(FULL, n == 15u);
		anyN |= nn != 0;
		if (lane == 0) {
			ws.plane[3][w] = hi;
			ws.plane[4][w] = lo;
			ws.plane[5][w] = nn;
		}
	}
	if (lane == 0) {
		for (int k = 0; k < 6; k++)
			ws.plane[k][k < 3 ? fwords : rwords] = 0;
	}
	__syncwarp();

	/* ---- K1: forward 8-mers into the hash (assembler.c:92-101) ---- */
	constexpr unsigned MASK = WS::SLOTS - 1;
	for (int base = 8; base < F; base += 32) {
		int p = base + lane;
		bool live = p < F;
		unsigned code = 0;
		if (live) {
			unsigned hi = window(ws.plane[0], p - 7) & 0xFFu;
			unsigned lo = window(ws.plane[1], p - 7) & 0xFFu;
			code = (hi << 8) | lo;
			if (anyN)
				live = (window(ws.plane[2], p - 8) & 0x1FFu) == 0;
		}
		unsigned entry = (code << 16) | (unsigned) p;
		unsigned slot = hash16(code, MASK);
		bool pending = live;
		while (__any_sync(FULL, pending)) {
			bool tryw = false;
			if (pending) {
				tryw = ws.htab[slot] == 0u;
				if (tryw)
					ws.htab[slot] = entry;
			}
			__syncwarp();
			if (pending) {
				if (tryw && ws.htab[slot] == entry)
					pending = false;
				else
					slot = (slot + 1) & MASK;
			}
			__syncwarp();
		}
	}
	__syncwarp();

	/* ---- K2: reverse 8-mers probe (assembler.c:104-110); flags = BIT_LIST_SET ---- */
	for (int base = 8; base < R; base += 32) {
		int e = base + lane;
		bool live = e < R;
		if (live) {
			unsigned hi = window(ws.plane[3], e - 7) & 0xFFu;
			unsigned lo = window(ws.plane[4], e - 7) & 0xFFu;
			unsigned code = (hi << 8) | lo;
			if (anyN)
				live = (window(ws.plane[5], e - 8) & 0x1FFu) == 0;
			if (live) {
				unsigned slot = hash16(code, MASK);
				unsigned m1 = 0xFFFFu, m2 = 0xFFFFu;
				for (;;) {
					unsigned ent = ws.htab[slot];
					if (ent == 0u)
						break;
					if ((ent >> 16) == code) {
						unsigned p = ent & 0xFFFFu;
						if (p < m1) {
							m2 = m1;
							m1 = p;
						} else if (p < m2) {
							m2 = p;
						}
					}
					slot = (slot + 1) & MASK;
				}
				if (m1 != 0xFFFFu) {
					int idx = F - (int) m1 + e - mo;
					if (idx >= 0 && idx < nbits)
						ws.cflag[idx] = 1;
				}
				if (m2 != 0xFFFFu) {
					int idx = F - (int) m2 + e - mo;
					if (idx >= 0 && idx < nbits)
						ws.cflag[idx] = 1;
				}
			}
		}
	}
	__syncwarp();
	/* ---- K3: clear the hash for the next pair (assembler.c:113-116) ---- */
	{
		uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *h4 = reinterpret_cast<uint4 *>(ws.htab);
		for (int k = lane; k < WS::SLOTS / 4; k += 32)
			h4[k] = z;
	}

	/* ---- K4/K5: sweep the candidates (assembler.c:118-143) ---- */
	const double qual_nn = prm->qual_nn;
	double best = qual_nn * (double) (unsigned long long) (F + R);   /* assembler.c:60 */
	int bestov = -1;
	int examined = 0;
	/* gather the flag bytes, 16 per lane, into a 16-bit mask per lane, and clear them */
	const int nflag_chunks = (nbits + 15) >> 4;
	bool none;
	{
		unsigned anyflag = 0;
		for (int c = lane; c < nflag_chunks; c += 32) {
			uint4 f = reinterpret_cast<const uint4 *>(ws.cflag)[c];
			anyflag |= f.x | f.y | f.z | f.w;
		}
		none = !__any_sync(FULL, anyflag != 0);
	}
	const int algo = prm->algo;
	for (int cbase = 0; cbase < nflag_chunks; cbase += 32) {
		int c = cbase + lane;
		unsigned m16 = 0;
		if (c < nflag_chunks) {
			uint4 f = reinterpret_cast<const uint4 *>(ws.cflag)[c];
			unsigned wv[4] = { f.x, f.y, f.z, f.w };
#pragma unroll
			for (int k = 0; k < 4; k++) {
#pragma unroll
				for (int b = 0; b < 4; b++)
					if ((wv[k] >> (8 * b)) & 0xFFu)
						m16 |= 1u << (k * 4 + b);
			}
			if (none)
				m16 = 0xFFFFu;
			/* only idx < nbits are candidates */
			int lim = nbits - c * 16;
			if (lim < 16)
				m16 &= (1u << lim) - 1u;
			reinterpret_cast<uint4 *>(ws.cflag)[c] = make_uint4(0, 0, 0, 0);
		}
		unsigned have;
		while ((have = __ballot_sync(FULL, m16 != 0)) != 0) {
			int leader = __ffs(have) - 1;
			unsigned lm = __shfl_sync(FULL, m16, leader);
			int bit = __ffs(lm) - 1;
			if (lane == leader)
				m16 &= m16 - 1;
			const int ov = (cbase + leader) * 16 + bit + mo;
			/* overlap_probability for this candidate, whole warp */
			const int i0 = max(0, ov - F), i1 = min(ov, R);    /* findex = F-ov+i in [0,F), template index i < R */
			double prob;
			if (algo == PB_SIMPLE_BAYES || algo == PB_FLASH) {
				int matches = 0, mism = 0, unk = 0;
				for (int i = i0 + lane; i < i1; i += 32) {
					unsigned f = nib(v.fnt, F - ov + i), r = nib(v.rnt, i);
					if (f == 15u || r == 15u)
						unk++;
					else if (f & r)
						matches++;
					else
						mism++;
				}
				unsigned packed = (unsigned) matches | ((unsigned) mism << 10) | ((unsigned) unk << 20);
				packed = __reduce_add_sync(FULL, packed);
				matches = packed & 1023;
				mism = (packed >> 10) & 1023;
				unk = packed >> 20;
				if (algo == PB_SIMPLE_BAYES) {
					/* algo_simple_bayes.c:61-65: size_t arithmetic inside the parenthesis */
					unsigned long long nn_count = (ov >= F && ov >= R)
						? (unsigned long long) unk
						: (unsigned long long) ((long long) F + R - 2 * (long long) ov + unk);
					prob = qual_nn * (double) nn_count + (double) matches * prm->sb_pmatch;
					prob = prob + (double) mism * prm->sb_pmismatch;
				} else {
					/* algo_flash.c:59: integer division inside log() */
					int real = matches + mism + unk, bad = mism + unk;
					prob = (real == 0) ? -2.0 : ((bad == real) ? 0.0 : -CUDART_INF);
				}
			} else {
				double acc = 0.0;
				for (int i = i0 + lane; i < i1; i += 32) {
					int fi = F - ov + i;
					unsigned f = nib(v.fnt, fi), r = nib(v.rnt, i);
					int qa = clampq(v.fq[fi]);
					if (algo == PB_PEAR) {
						/* algo_pear.c:52,54 index the FORWARD qualities with rindex = R-1-i; past the
						 * end of the forward read that is defined as quality 0. */
						int ri = R - 1 - i;
						int qb = (ri < F) ? clampq(v.fq[ri]) : 0;
						if (f == 15u || r == 15u)
							acc -= prm->pear_random_base;
						else
							acc += s_over[(((f & r) ? 1 : 0) * PB_NQ + qa) * PB_NQ + qb];
					} else {
						int qb = clampq(v.rq[i]);
						acc += s_over[(((f & r) ? 1 : 0) * PB_NQ + qa) * PB_NQ + qb];
					}
				}
				prob = warp_sum(acc);
			}
			if (prob > best) {        /* strict, ascending overlap: assembler.c:128-131 */
				best = prob;
				bestov = ov;
			}
			examined++;
		}
	}
	res.examined = (uint16_t) examined;
	if ((long long) examined == (long long) maxov - mo + 1)    /* assembler.c:135-137 */
		res.slow = 1;
	if (bestov < 0) {
		res.status = PB_PAIR_NOALGN;
		return;
	}
	/* ---- K6: reconstruction (assembler.c:145-250) ---- */
	const int len = F - fo - bestov + R - ro + 1;
	if (len <= 0 || len > 2 * PB_MAX_LEN) {
		res.status = PB_PAIR_NOALGN;
		return;
	}
	const int seq_len = len - 1;
	const int df = F - fo - bestov, dr = R - ro - bestov;
	const int dfp = max(df, 0), dfn = min(df, 0), drn = min(dr, 0);
	const int nover = bestov + dfn + drn;
	/* B-cliff: trailing run of '#'/qual 2 in each read (assembler.c:176-177) */
	int unmasked_f = F, lead_r = 0;
	for (int base = 0; base < F; base += 32) {
		int i = F - 1 - base - lane;
		unsigned notb = __ballot_sync(FULL, !(i >= 0 && v.fq[i] == 2));
		if (notb) {
			unmasked_f = F - base - (__ffs(notb) - 1);
			break;
		}
		unmasked_f = max(F - base - 32, 0);
	}
	for (int base = 0; base < R; base += 32) {
		int j = base + lane;      /* template order: reverse[R-1-j] */
		unsigned notb = __ballot_sync(FULL, !(j < R && v.rq[j] == 2));
		if (notb) {
			lead_r = base + (__ffs(notb) - 1);
			break;
		}
		lead_r = min(base + 32, R);
	}
	double qsum = 0.0;
	int mism = 0, degen = 0;
	for (int base = 0; base < seq_len; base += 32) {
		int idx = base + lane;
		if (idx < seq_len) {
			unsigned nt;
			int a, b, m = 0;
			if (idx < dfp) {                       /* forward only, assembler.c:162-173 */
				int fi = idx + fo;
				nt = nib(v.fnt, fi);
				a = clampq(v.fq[fi]);
				b = PB_NQ;
			} else if (idx < dfp + nover) {        /* overlap, assembler.c:181-228 */
				int i = idx - dfp;
				int fi = fo + dfp + i;
				int j = i - dfn;                   /* template index of reverse[R-1-i+dfn] */
				unsigned fn = nib(v.fnt, fi), rn = nib(v.rnt, j);
				int fqv = v.fq[fi], rqv = v.rq[j];
				m = (fn & rn) ? 1 : 0;
				if (!m)
					mism++;
				a = (fi >= unmasked_f) ? PB_NQ : clampq(fqv);
				b = (j < lead_r) ? PB_NQ : clampq(rqv);
				nt = m ? (fn & rn) : ((fqv < rqv) ? rn : fn);
			} else {                               /* reverse only, assembler.c:231-243 */
				int j = bestov + (idx - dfp - nover);
				nt = nib(v.rnt, j);
				a = PB_NQ;
				b = clampq(v.rq[j]);
			}
			double p = s_recon[(m * PB_NQM + a) * PB_NQM + b];
			qsum += p;
			if (__popc(nt) != 1)
				degen++;
			if (out_nt && idx < out_cap)
				out_nt[idx] = (uint8_t) nt;
			if (out_p && idx < out_cap)
				out_p[idx] = p;
		}
	}
	qsum = warp_sum(qsum);
	mism = warp_sum_int(mism);
	degen = warp_sum_int(degen);
	res.quality = qsum / (double) len;               /* assembler.c:244: divides by len, not seq_len */
	res.overlap = (uint16_t) bestov;
	res.est_prob = best;
	res.seq_len = (uint16_t) seq_len;
	res.mismatches = (uint16_t) mism;
	res.degenerates = (uint16_t) degen;
	if (res.quality < prm->threshold)                /* assembler.c:334-338 */
		res.status = PB_PAIR_LOWQ;
}

template <int ML, bool OVER, int WARPS_PER_BLOCK>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
assemble_kernel(const pb_device_params *__restrict__ prm, int n,
                const uint8_t *__restrict__ reads, const pb_pair_meta *__restrict__ meta,
                pb_pair_result *__restrict__ results, uint8_t *__restrict__ seq_nt, double *__restrict__ seq_p,
                long long seq_stride, unsigned long long *__restrict__ counters) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	using WS = WarpSmem<ML>;
	/* block-level: LUTs + counters, then the per-warp areas */
	constexpr int OVER_N = OVER ? 2 * PB_NQ * PB_NQ : 0;
	double *s_recon = reinterpret_cast<double *>(smem_raw);
	double *s_over = s_recon + 2 * PB_NQM * PB_NQM;
	double *s_score = s_over + OVER_N;
	double *s_score_err = s_score + PB_NQM;
	unsigned *s_cnt = reinterpret_cast<unsigned *>(s_score_err + PB_NQM);
	constexpr size_t LUT_BYTES = (2 * PB_NQM * PB_NQM + OVER_N + 2 * PB_NQM) * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned);
	constexpr size_t LUT_ALIGNED = (LUT_BYTES + 127) & ~(size_t) 127;
	WS *wsall = reinterpret_cast<WS *>(smem_raw + LUT_ALIGNED);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < 2 * PB_NQM * PB_NQM; i += blockDim.x)
		s_recon[i] = (&prm->recon[0][0][0])[i];
	if (OVER)
		for (int i = tid; i < 2 * PB_NQ * PB_NQ; i += blockDim.x)
			s_over[i] = (&prm->over[0][0][0])[i];
	for (int i = tid; i < PB_NQM; i += blockDim.x) {
		s_score[i] = prm->score[i];
		s_score_err[i] = prm->score_err[i];
	}
	for (int i = tid; i < PB_NCOUNTERS; i += blockDim.x)
		s_cnt[i] = 0;
	WS &ws = wsall[warp];
	for (int k = lane; k < WS::SLOTS; k += 32)
		ws.htab[k] = 0;
	for (int k = lane; k < WS::NFLAG; k += 32)
		ws.cflag[k] = 0;
	if (lane == 0) {
		for (int s = 0; s < NSTAGE; s++)
			mbar_init(&ws.bar[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const int wglobal = blockIdx.x * WARPS_PER_BLOCK + warp;
	const int wstride = gridDim.x * WARPS_PER_BLOCK;

	auto issue = [&](int pair, int stage) {
		if (lane == 0) {
			pb_pair_meta m = meta[pair];
			ws.meta[stage] = m;
			unsigned bytes = record_bytes(m.flen, m.rlen);
			mbar_expect_tx(&ws.bar[stage], bytes);
			if (bytes)
				bulk_g2s(ws.stage[stage], reads + (size_t) m.off16 * 16, bytes, &ws.bar[stage]);
		}
	};

	int pair = wglobal;
	if (pair < n)
		issue(pair, 0);
	for (int it = 0; pair < n; it++, pair += wstride) {
		const int stage = it & 1;
		const int next = pair + wstride;
		if (next < n)
			issue(next, stage ^ 1);
		mbar_wait(&ws.bar[stage], (it >> 1) & 1);
		const pb_pair_meta m = ws.meta[stage];
		union { pb_pair_result r; uint4 v[2]; } ru;
		pb_pair_result &res = ru.r;
		uint8_t *o_nt = seq_nt ? seq_nt + (size_t) pair * seq_stride : nullptr;
		double *o_p = seq_p ? seq_p + (size_t) pair * seq_stride : nullptr;
		process_pair<ML>(ws, ws.stage[stage], m.flen, m.rlen, prm, s_recon, s_over, s_score, s_score_err, res, o_nt, o_p, (int) seq_stride, lane);
		if (lane == 0) {
			uint4 *dst = reinterpret_cast<uint4 *>(&results[pair]);
			dst[0] = ru.v[0];
			dst[1] = ru.v[1];
			atomicAdd(&s_cnt[PB_C_COUNT], 1u);
			if (res.slow)
				atomicAdd(&s_cnt[PB_C_SLOW], 1u);
			switch (res.status) {
			case PB_PAIR_OK:
				atomicAdd(&s_cnt[PB_C_OK], 1u);
				atomicAdd(&s_cnt[PB_C_OVERLAPS + res.overlap], 1u);
				atomicMax(&s_cnt[PB_C_LONGEST], (unsigned) res.overlap);
				break;
			case PB_PAIR_LOWQ: atomicAdd(&s_cnt[PB_C_LOWQ], 1u); break;
			case PB_PAIR_NOALGN: atomicAdd(&s_cnt[PB_C_NOALGN], 1u); break;
			case PB_PAIR_BADR: atomicAdd(&s_cnt[PB_C_BADR], 1u); break;
			case PB_PAIR_NOFP: atomicAdd(&s_cnt[PB_C_NOFP], 1u); break;
			case PB_PAIR_NORP: atomicAdd(&s_cnt[PB_C_NORP], 1u); break;
			}
		}
		__syncwarp();      /* every lane is done with this stage before it is refilled */
	}
	__syncthreads();
	for (int i = tid; i < PB_NCOUNTERS; i += blockDim.x) {
		unsigned c = s_cnt[i];
		if (c) {
			if (i == PB_C_LONGEST)
				atomicMax(&counters[i], (unsigned long long) c);
			else
				atomicAdd(&counters[i], (unsigned long long) c);
		}
	}
}

template <int ML, bool OVER, int WARPS_PER_BLOCK> constexpr size_t assemble_smem_bytes() {
	constexpr size_t LUT_BYTES = (2 * PB_NQM * PB_NQM + (OVER ? 2 * PB_NQ * PB_NQ : 0) + 2 * PB_NQM) * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned);
	constexpr size_t LUT_ALIGNED = (LUT_BYTES + 127) & ~(size_t) 127;
	return LUT_ALIGNED + sizeof(WarpSmem<ML>) * WARPS_PER_BLOCK;
}

/* ---- pack: flat AoS panda_qual -> packed records (one warp per pair) -------------- */
__global__ void pack_kernel(int n, const uint8_t *__restrict__ f_data, const unsigned long long *__restrict__ f_off,
                            const uint8_t *__restrict__ r_data, const unsigned long long *__restrict__ r_off,
                            const uint32_t *__restrict__ rec_off16, uint8_t *__restrict__ reads, pb_pair_meta *__restrict__ meta) {
	const int lane = threadIdx.x & 31;
	const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (pair >= n)
		return;
	const unsigned long long fb = f_off[pair], rb = r_off[pair];
	const int F = (int) (f_off[pair + 1] - fb), R = (int) (r_off[pair + 1] - rb);
	const uint8_t *f = f_data + 2 * fb, *r = r_data + 2 * rb;     /* {nt, qual} byte pairs */
	uint8_t *rec = reads + (size_t) rec_off16[pair] * 16;
	const int fw = ((F + 7) / 8) * 4, rw = ((R + 7) / 8) * 4, fqb = ((F + 3) / 4) * 4, rqb = ((R + 3) / 4) * 4;
	if (lane == 0) {
		pb_pair_meta m;
		m.off16 = rec_off16[pair];
		m.flen = (uint16_t) F;
		m.rlen = (uint16_t) R;
		meta[pair] = m;
	}
	for (int k = lane; k < fw; k += 32) {            /* forward nibbles */
		int b0 = 2 * k, b1 = 2 * k + 1;
		unsigned lo = b0 < F ? (f[2 * b0] & 15u) : 0u, hi = b1 < F ? (f[2 * b1] & 15u) : 0u;
		rec[k] = (uint8_t) (lo | (hi << 4));
	}
	for (int k = lane; k < rw; k += 32) {            /* reverse nibbles, template order */
		int b0 = 2 * k, b1 = 2 * k + 1;
		unsigned lo = b0 < R ? (r[2 * (R - 1 - b0)] & 15u) : 0u, hi = b1 < R ? (r[2 * (R - 1 - b1)] & 15u) : 0u;
		rec[fw + k] = (uint8_t) (lo | (hi << 4));
	}
	for (int k = lane; k < fqb; k += 32)
		rec[fw + rw + k] = k < F ? f[2 * k + 1] : 0;
	for (int k = lane; k < rqb; k += 32)
		rec[fw + rw + fqb + k] = k < R ? r[2 * (R - 1 - k) + 1] : 0;
	const int used = fw + rw + fqb + rqb, total = (used + 15) & ~15;
	for (int k = used + lane; k < total; k += 32)
		rec[k] = 0;
}

}  // namespace pb
