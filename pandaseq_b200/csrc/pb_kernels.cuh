/* pb_kernels.cuh -- sm_100a kernels for the PANDAseq pair-assembly hot path (kernel v2).
 *
 * assemble_kernel<ML>: one warp owns one read pair at a time (persistent grid, warps
 * stride over the batch).  Per pair, following the reference's align()
 * (assembler.c:48-250) and assemble_seq() (assembler.c:252-348):
 *
 *   stage   the packed record (4-bit nt + 8-bit PHRED, both reads) is pulled from HBM
 *           into this warp's shared-memory stage by ONE bulk async copy (cp.async.bulk,
 *           the 1-D TMA path; SASS UBLKCP) completing on an mbarrier; the copy for the
 *           warp's next pair is in flight while the current pair is processed.
 *   primers panda_compute_offset_qual (offset.c:47-112), when primers are set.
 *   codes   a lane takes one 32-bit word (8 bases) of a read, turns the nibbles into
 *           2-bit k-mer digits with bit-parallel logic (misc.h:41: T=3 G=2 C=1 else 0),
 *           and emits the 16-bit code of the eight 8-mers that end in its word.
 *   seed    K1-K3 of align(): forward 8-mers go into a per-warp bucket table in shared
 *           memory (4 x 16-bit entries per bucket; same-bucket lanes of a round are
 *           ranked with match.any so every round is a fixed instruction sequence and
 *           entries stay in position order); reverse 8-mers read one bucket and keep
 *           the first two forward positions with their code -- the observable behaviour
 *           of the reference's 65536x2 table (SURVEY.md §8a "table-free statement").
 *           A bucket that would need a 5th entry (low-complexity reads) sends the pair
 *           through an exact open-addressing path in the same memory.
 *   score   K4/K5: every flagged overlap (or all, if none) is scored, a lane per 8 bases.
 *   merge   K6: merged read (8 bases per lane, bit-parallel), per-base log p from a
 *           2x48x48 LUT in shared memory, quality = sum / len, counts; result record.
 *
 * Floating point: compiled with --fmad=false.  simple_bayes / flash scores are closed
 * forms of integer counts and are bit-identical to the reference.  The pear / rdp_mle
 * score and the quality sum are sums of LUT entries; the reference adds them left to
 * right, the warp adds per-lane partial sums with a shuffle tree: |difference| ~1e-13,
 * tolerance 1e-6 (BASELINE.json north_star).
 */
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "pb_internal.h"

namespace pb {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int NSTAGE = 2;
constexpr unsigned NIB1 = 0x11111111u;

/* ---- per-warp shared memory layout, sized by the read-length class ML -------- */
template <int ML> struct WarpSmem {
	static constexpr int NTW = (ML + 7) / 8;                 /* nibble words per read */
	static constexpr int STAGE_BYTES = ((2 * NTW * 4 + 2 * ((ML + 3) / 4) * 4) + 15) & ~15;
	static constexpr int PBITS = (ML <= 256) ? 8 : 9;        /* bits of a forward position */
	static constexpr int IDXBITS = (ML <= 160) ? 9 : (ML <= 320 ? 10 : 11);
	static constexpr int NB = 1 << IDXBITS;                  /* buckets of 4 entries */
	static constexpr int TAGBITS = 16 - IDXBITS;
	static_assert(TAGBITS + PBITS <= 16, "bucket entry must fit 16 bits");
	static constexpr int SLOTS = NB * 2;                     /* the same memory as 32-bit open-addressing slots */
	static constexpr int CODEN = NTW * 8 + 8;
	static constexpr int NFLAG = ((2 * ML + 15) & ~15) + 16;
	alignas(128) uint8_t stage[NSTAGE][STAGE_BYTES];
	alignas(16) uint64_t btab[NB];
	alignas(16) uint32_t bcnt[NB / 8];                       /* entries per bucket, 4 bits each */
	alignas(16) uint16_t code_f[CODEN];
	alignas(16) uint16_t code_r[CODEN];
	alignas(16) uint8_t inval_f[(NTW + 19) & ~15];           /* bit t of byte w: k-mer ending at 8w+t is invalid */
	alignas(16) uint8_t inval_r[(NTW + 19) & ~15];
	alignas(16) uint8_t cflag[NFLAG];
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
__device__ __forceinline__ unsigned lanemask_lt() {
	unsigned m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}
/* 8 nibbles starting at base `pos` of a nibble array; pos may be negative or run past the read (the
 * caller masks those nibbles; the loads stay inside this warp's shared memory / the LUT area before it). */
__device__ __forceinline__ unsigned nibwin(const uint32_t *w32, int pos) {
	int j = max(pos >> 3, -1);
	return __funnelshift_r(w32[j], w32[j + 1], (pos & 7) * 4);
}
/* low n nibbles set, n in [0, 8] */
__device__ __forceinline__ unsigned nibmask(int n) {
	return __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - 4 * n);
}
/* bit 4k set iff nibble k is non-zero */
__device__ __forceinline__ unsigned nz_nib(unsigned x) {
	unsigned t = x | (x >> 1);
	return (t | (t >> 2)) & NIB1;
}
/* bit 4k set iff nibble k == 15 (N) */
__device__ __forceinline__ unsigned n_nib(unsigned x) {
	unsigned t = x & (x >> 1);
	return t & (t >> 2) & NIB1;
}
/* bit 4k set iff nibble k has exactly one bit set */
__device__ __forceinline__ unsigned onehot_nib(unsigned x) {
	unsigned x1 = x >> 1, x2 = x >> 2, x3 = x >> 3;
	unsigned par = x ^ x1 ^ x2 ^ x3;
	unsigned two = (x & x1) | (x2 & x3) | ((x | x1) & (x2 | x3));
	return par & ~two & NIB1;
}
/* nibble-spaced bits (bit 4k) -> 8 contiguous bits */
__device__ __forceinline__ unsigned squeeze1(unsigned v) {
	unsigned t = (v | (v >> 3)) & 0x03030303u;
	t = (t | (t >> 6)) & 0x000F000Fu;
	return (t | (t >> 12)) & 0xFFu;
}
/* nibble-spaced 2-bit fields (bits 4k, 4k+1) -> 16 contiguous bits */
__device__ __forceinline__ unsigned squeeze2(unsigned v) {
	unsigned t = (v | (v >> 2)) & 0x0F0F0F0Fu;
	t = (t | (t >> 4)) & 0x00FF00FFu;
	return (t | (t >> 8)) & 0xFFFFu;
}
__device__ __forceinline__ unsigned hash_slot(unsigned code, unsigned mask) {
	return ((code * 40503u) >> 3) & mask;
}
/* same arithmetic as pb_record_bytes() in the public header */
__device__ __forceinline__ unsigned record_bytes(unsigned flen, unsigned rlen) {
	unsigned b = ((flen + 7) / 8) * 4 + ((rlen + 7) / 8) * 4 + ((flen + 3) / 4) * 4 + ((rlen + 3) / 4) * 4;
	return (b + 15u) & ~15u;
}

/* plugin_pear_test.c:18-39 on the result record; cdf = pb_device_params.pear_cdf (the plugin's inner sums, built on the host
 * with its own libm calls).  A limit that is negative or not a number makes the plugin's loop run for 2^64 iterations; by
 * then it has added all i + 1 non-zero terms, which is what cdf[i][i + 1] holds. */
__device__ __forceinline__ bool pear_test_pass(const double *__restrict__ cdf, double alpha, double beta, double cutoff,
                                               int overlap, int mismatches, int F, int R) {
	double product = 1.0;
	const double oes = alpha * (double) (unsigned long long) (overlap - mismatches) + beta * (double) (unsigned long long) mismatches;
	const int lim = min(F, R);
	for (int i = overlap; i < lim; i++) {
		const double lraw = ceil((oes - beta * (double) i) / (alpha - beta)) - 1.0;
		const int l = (lraw >= 0.0 && lraw < (double) (i + 1)) ? (int) lraw : i + 1;
		product *= cdf[(size_t) i * PB_PEAR_COLS + l];
	}
	return cutoff > 1.0 - product * product;
}

struct PairView {
	const uint8_t *fnt, *rnt;      /* packed nibbles; rnt in template order */
	const uint32_t *fnt32, *rnt32;
	const int8_t *fq, *rq;         /* raw PHRED chars; rq in template order */
	int F, R;
};

/* offset.c:47-112 for one read, whole warp.  Each START offset s is owned by a lane: the alignment that
 * begins at read position s accumulates primer[x] vs read[s+x] for x = 0..P-1 in that order, exactly the
 * order the reference's circular buffer receives its addends, so the sum is bit-identical.
 *
 * TEMPLATE_ORDER: the read is stored reversed (reverse read), so read position i is element len-1-i.
 * Returns bestindex as the reference does (0 = not found, else 1 + bases consumed).
 * With penalty == 0 the comparison exp(a) > exp(b) is done as a > b (exp is monotone; SURVEY.md §8a a18
 * measured 0 differences on 600 k reads); with a penalty, CUDA's exp() is used.
 * score[] and score_err[] are adjacent 48-entry tables; qoff[256 + raw] = clamp(raw) * 8. */
template <bool TEMPLATE_ORDER>
__device__ int primer_offset(const uint32_t *nt32, const int8_t *q, int len,
                             const uint8_t *primer, int P, double threshold, double penalty,
                             const double *score, const uint16_t *qoff, double *se, int lane) {
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
	const char *tab = reinterpret_cast<const char *>(score);
	const unsigned char *qu = reinterpret_cast<const unsigned char *>(q);
	/* Both scores of every read position, once: se[2i] = qual_score[q_i] (the base agrees with the primer),
	 * se[2i+1] = qual_score_err[q_i] (offset.c:93-101), i in scan order.  The scan below then costs one predicate,
	 * one select and one 8-byte load per (start, primer base). */
	auto stage_scores = [&]() {
		for (int i = lane; i < len; i += 32) {
			const unsigned off = qoff[256 + qu[TEMPLATE_ORDER ? (len - 1 - i) : i]];
			double2 v;
			v.x = *reinterpret_cast<const double *>(tab + off);
			v.y = *reinterpret_cast<const double *>(tab + off + PB_NQM * 8);      /* score_err[] follows score[] */
			reinterpret_cast<double2 *>(se)[i] = v;
		}
		__syncwarp();
	};
	/* The primer as masks: pm[x] = its nibble at the place base x & 7 of an 8-base window has (template order: the window
	 * is bit-reversed below, so the nibble is too); all ones for N, which contributes nothing (offset.c:97). */
	uint32_t *pm = reinterpret_cast<uint32_t *>(se + 2 * len);
	bool has_n = false;
	for (int x = lane; x - lane < P; x += 32) {
		unsigned pn = x < P ? primer[x] : 0u;
		const bool isn = pn == 15u;
		has_n = has_n || isn;
		if (TEMPLATE_ORDER)
			pn = __brev(pn) >> 28;
		if (x < P)
			pm[x] = isn ? ~0u : pn << (4 * (x & 7));
	}
	has_n = __any_sync(FULL, has_n);
	__syncwarp();
	/* the sum of start s, every term in primer order */
	auto exact_sum = [&](int s) -> double {
		double sum = 0.0;
		const int el = TEMPLATE_ORDER ? (len - 1 - s) : s;   /* element holding read position s */
		const char *sp = reinterpret_cast<const char *>(se + 2 * s);
		for (int x0 = 0; x0 < P; x0 += 8) {
			/* nibbles of the 8 read positions s+x0 .. s+x0+7 (template order: element first+7 is position s+x0) */
			unsigned w = nibwin(nt32, TEMPLATE_ORDER ? (el - x0 - 7) : (el + x0));
			if (TEMPLATE_ORDER)
				w = __brev(w);                         /* nibble k of the reversed word holds position s+x0+k, bits reversed inside it */
			const uint32_t *pmb = pm + x0;
			const char *spb = sp + 16 * x0;
			if (P - x0 >= 8 && !has_n) {               /* warp-uniform: a full block of a primer without N */
#pragma unroll
				for (int t = 0; t < 8; t++) {
					const unsigned off = (w & pmb[t]) == 0u ? 8u : 0u;
					sum += *reinterpret_cast<const double *>(spb + 16 * t + off);
				}
			} else {
				const int xn = min(P - x0, 8);
				for (int t = 0; t < xn; t++) {
					const unsigned m = pmb[t];
					if (m != ~0u) {                        /* warp-uniform: not an N of the primer */
						const unsigned off = (w & m) == 0u ? 8u : 0u;
						sum += *reinterpret_cast<const double *>(spb + 16 * t + off);
					}
				}
			}
		}
		return sum;
	};
	if (penalty == 0.0 && nstart > 0) {
		/* Without a penalty the winner is the start with the largest sum / (index + 1), the lowest one among equals, if that beats
		 * P * threshold.  Every term is a log-probability, i.e. <= 0, and a base that disagrees with the primer adds
		 * qual_score_err[q] <= emax, the largest such value over this read's qualities.  So a start with m disagreeing bases among
		 * the primer's first eight cannot reach more than m * emax / (index + 1): a first pass counts those m for every start (one
		 * AND + POPC on the window instead of P loads and additions), the start with the fewest is summed exactly, and after that
		 * only starts whose bound still reaches the best value are.  On a read that carries the primer the bound removes every other
		 * start; the sums that are formed are formed exactly as before, so the result is the full scan's, bit for bit. */
		/* qual_score_err[] falls with the quality (mktable.c:75-82: log of the error probability), so emax belongs to the lowest
		 * quality of the read (as a signed char: PHREDCLAMP takes anything below zero to zero) */
		int qmin = 127;
		for (int i = lane; i < len; i += 32)
			qmin = min(qmin, (int) q[i]);
		qmin = __reduce_min_sync(FULL, qmin);
		const double emax = *reinterpret_cast<const double *>(tab + qoff[256 + (unsigned) (unsigned char) qmin] + PB_NQM * 8);
		/* the primer's first eight bases as one word of nibbles, and how many of them count: an N adds nothing (offset.c:97) */
		const unsigned pmine = (lane < 8 && lane < P) ? pm[lane] : ~0u;
		const unsigned pw0 = __reduce_or_sync(FULL, pmine == ~0u ? 0u : pmine);
		const unsigned np0 = (unsigned) __popc(__ballot_sync(FULL, pmine != ~0u));
		/* m of this lane's starts, four bits per round (15 where the lane has no start), rounds 0..7 in mlo, 8..15 in mhi */
		unsigned mlo = 0xFFFFFFFFu, mhi = 0xFFFFFFFFu, keymin = 0x7FFFFFFFu;
		auto count_round = [&](int base) -> unsigned {
			const int s = min(base + lane, nstart - 1);
			const int el = TEMPLATE_ORDER ? (len - 1 - s) : s;
			unsigned w = nibwin(nt32, TEMPLATE_ORDER ? (el - 7) : el);
			if (TEMPLATE_ORDER)
				w = __brev(w);
			unsigned m = np0 - (unsigned) __popc(nz_nib(w & pw0));      /* the window's nibbles outside the primer meet zeros */
			m = base + lane < nstart ? m : 15u;
			keymin = min(keymin, (m << 16) | (unsigned) s);
			return m;
		};
		{
			unsigned acc = 0;
			int sh = 0;
			for (int base = 0; base < nstart && sh < 32; base += 32, sh += 4)
				acc |= count_round(base) << sh;
			mlo = acc | (sh < 32 ? 0xFFFFFFFFu << sh : 0u);
			acc = 0;
			sh = 0;
			for (int base = 256; base < nstart; base += 32, sh += 4)
				acc |= count_round(base) << sh;
			mhi = acc | (sh < 32 ? 0xFFFFFFFFu << sh : 0u);
		}
		keymin = __reduce_min_sync(FULL, keymin);
		/* starts that agree with all of the primer's first eight bases: the zero nibbles */
		auto zero_nibbles = [](unsigned x) { x |= x >> 1; x |= x >> 2; return (unsigned) __popc(~x & NIB1); };
		const unsigned clean = __reduce_add_sync(FULL, zero_nibbles(mlo) + zero_nibbles(mhi));
		const int s0 = (int) (keymin & 0xFFFFu);
		int best_s = -1;                               /* -1: `best` is still the threshold, which an equal value does not beat */
		{
			/* The sum of start s0 straight from the tables: the same addends in the same order as exact_sum(), which needs both
			 * scores of every read position staged first -- not worth it for one start.  Lane x fetches the term of primer base x,
			 * the terms are then added in primer order. */
			const uint8_t *ntb = reinterpret_cast<const uint8_t *>(nt32);
			double sum = 0.0;
			for (int xb = 0; xb < P; xb += 32) {
				const int x = xb + lane;
				double term = 0.0;
				bool use = false;
				if (x < P && primer[x] != 15u) {
					const int pos = s0 + x, el = TEMPLATE_ORDER ? (len - 1 - pos) : pos;
					const bool hit = (nib(ntb, el) & primer[x]) != 0u;
					term = *reinterpret_cast<const double *>(tab + qoff[256 + qu[el]] + (hit ? 0u : PB_NQM * 8u));
					use = true;
				}
				const unsigned um = __ballot_sync(FULL, use);
				const int xn = min(P - xb, 32);
				if (um == (xn == 32 ? 0xFFFFFFFFu : (1u << xn) - 1u)) {      /* no N among these primer bases */
					for (int t = 0; t < xn; t++)
						sum += __shfl_sync(FULL, term, t);
				} else {
					for (int t = 0; t < xn; t++) {
						const double v = __shfl_sync(FULL, term, t);
						if ((um >> t) & 1u)
							sum += v;
					}
				}
			}
			const double v0 = sum / (double) (s0 + P + 1);
			if (v0 > best) {
				best = v0;
				best_s = s0;
			}
		}
		/* One disagreeing base already rules a start out if it does so at the largest index, len; then only the other clean starts
		 * are left to look at -- on a read that carries the primer once, none. */
		if (emax * 0.999999999 < best * (double) len && clean == ((keymin >> 16) == 0u ? 1u : 0u)) {
			__syncwarp();
			return best_s < 0 ? 0 : best_s + P + 1;
		}
		stage_scores();
		int rnd = 0;
		for (int base = 0; base < nstart; base += 32, rnd++) {
			const int s = base + lane;
			const unsigned m = ((rnd < 8 ? mlo >> (4 * rnd) : mhi >> (4 * (rnd - 8))) & 15u);
			/* skipped only if the bound, loosened by 1e-9 of itself against the rounding of either side, stays below the best */
			const bool cand = s < nstart && s != s0 && !((double) m * emax * 0.999999999 < best * (double) (s + P + 1));
			if (!__any_sync(FULL, cand))
				continue;
			double val = -CUDART_INF;
			if (cand)
				val = exact_sum(s) / (double) (s + P + 1);
			int who = s;
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				const double ov = __shfl_xor_sync(FULL, val, d);
				const int ow = __shfl_xor_sync(FULL, who, d);
				if (ov > val || (ov == val && ow < who)) {
					val = ov;
					who = ow;
				}
			}
			if (val > best || (val == best && best_s >= 0 && who < best_s)) {
				best = val;
				best_s = who;
			}
		}
		__syncwarp();      /* se[] is rewritten by the next scan */
		return best_s < 0 ? 0 : best_s + P + 1;
	}
	stage_scores();
	for (int base = 0; base < nstart; base += 32) {
		const int s = min(base + lane, nstart - 1);    /* surplus lanes redo the last start; masked below */
		const double sum = exact_sum(s);
		const int index = s + P;                   /* the index at which this slot is examined */
		double val = sum / (double) (index + 1);
		if (penalty != 0.0)
			val = exp(val) - (double) index * penalty;
		if (base + lane >= nstart)
			val = -CUDART_INF;
		/* The reference scans starts in increasing order and keeps the first strictly better one:
		 * within a batch that is the maximum value with the lowest start on ties. */
		if (__any_sync(FULL, val > best)) {
			int who = s;
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				const double ov = __shfl_xor_sync(FULL, val, d);
				const int ow = __shfl_xor_sync(FULL, who, d);
				if (ov > val || (ov == val && ow < who)) {
					val = ov;
					who = ow;
				}
			}
			best = val;
			best_index = who + P + 1;
		}
	}
	__syncwarp();      /* se[] is rewritten by the next scan */
	return best_index;
}

/* offset.c:47-90 with the result_base_score scorer (offset.c:114-133), for the primers-after path: the haystack is
 * the assembled sequence (4-bit bases `snt32`, per-base log p `sp`, both in this warp's global scratch).
 * For p < 0 the reference's "not p" is log(-expm1(-p)) = NaN, so a start offset with any mismatching base can
 * never win (NaN > x is false): such starts are dropped, the others sum p in primer order. */
__device__ int primer_offset_result(const uint32_t *snt32, const double *sp, int len, bool reverse,
                                    const uint8_t *primer, int P, double threshold, double penalty, int lane) {
	if (P > len)
		return 0;
	double best = (double) P * threshold;
	if (penalty != 0.0)
		best = exp(best);
	int best_index = 0;
	const int nstart = len - P;
	const uint8_t *snt = reinterpret_cast<const uint8_t *>(snt32);
	for (int base = 0; base < nstart; base += 32) {
		const int s = min(base + lane, nstart - 1);
		double sum = 0.0;
		bool dead = false;
		for (int x = 0; x < P; x++) {
			const unsigned pn = primer[x];
			if (pn == 15u)
				continue;
			const int el = reverse ? (len - 1 - (s + x)) : (s + x);
			if (nib(snt, el) & pn)
				sum += sp[el];
			else
				dead = true;
		}
		const int index = s + P;
		double val = sum / (double) (index + 1);
		if (penalty != 0.0)
			val = exp(val) - (double) index * penalty;
		if (dead || base + lane >= nstart)
			val = -CUDART_INF;
		int who = s;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			const double ov = __shfl_xor_sync(FULL, val, d);
			const int ow = __shfl_xor_sync(FULL, who, d);
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

/* ---- k-mer codes ---------------------------------------------------------------------------------
 * For every position p of a read, the 16-bit code of the 8-mer ending at p (digits of misc.h:41, oldest
 * base in the low bits -- any fixed injective packing will do, both reads use the same one).
 * Returns per-lane flags: bit 0 = some nibble is N, bit 1 = some nibble inside the read is not A/C/G/T. */
template <int NTW>
__device__ __forceinline__ unsigned gen_codes(const uint32_t *nt32, int len, uint16_t *codes, int lane) {
	const int nw = (len + 7) >> 3;
	unsigned flags = 0, carry = 0;
#pragma unroll
	for (int w0 = 0; w0 < NTW; w0 += 32) {
		if (w0 < nw) {               /* warp-uniform */
			const int w = w0 + lane;
			const unsigned x = (w < nw) ? nt32[w] : 0u;
			const unsigned x1 = x >> 1, x2 = x >> 2, x3 = x >> 3;
			const unsigned lo = (x3 ^ x1) & ~x2 & ~x & NIB1;      /* T or C */
			const unsigned hi = (x3 ^ x2) & ~x1 & ~x & NIB1;      /* T or G */
			const unsigned isa = x & ~x1 & ~x2 & ~x3 & NIB1;
			const unsigned inread = nibmask(min(max(len - 8 * w, 0), 8)) & NIB1;
			if ((x & x1 & x2 & x3 & NIB1) != 0)
				flags |= 1u;
			if ((inread & ~(lo | hi | isa)) != 0)
				flags |= 2u;
			const unsigned c16 = squeeze2(lo | (hi << 1));
			unsigned prev = __shfl_up_sync(FULL, c16, 1);
			if (lane == 0)
				prev = carry;
			carry = __shfl_sync(FULL, c16, 31);
			const unsigned W = prev | (c16 << 16);    /* bases 8w-8 .. 8w+7, two bits each */
			if (w < nw) {
				uint4 o;
				o.x = ((W >> 2) & 0xFFFFu) | ((W >> 4) << 16);
				o.y = ((W >> 6) & 0xFFFFu) | ((W >> 8) << 16);
				o.z = ((W >> 10) & 0xFFFFu) | ((W >> 12) << 16);
				o.w = ((W >> 14) & 0xFFFFu) | ((W >> 16) << 16);
				reinterpret_cast<uint4 *>(codes)[w] = o;
			}
		}
	}
	return flags;
}

/* Rare path (a read contains N): bit t of inval[w] is set when the 9 bases ending at 8w+t contain an N
 * or 8w+t < 8 -- the `bad` counter of misc.h:41. */
template <int NTW>
__device__ __noinline__ void gen_invalid(const uint32_t *nt32, int len, uint8_t *inval, int lane) {
	const int nw = (len + 7) >> 3;
	unsigned carry = 0;
	for (int w0 = 0; w0 < nw; w0 += 32) {
		const int w = w0 + lane;
		const unsigned x = (w < nw) ? nt32[w] : 0u;
		const unsigned n8 = squeeze1(n_nib(x));
		unsigned prev = __shfl_up_sync(FULL, n8, 1);
		if (lane == 0)
			prev = carry;
		carry = __shfl_sync(FULL, n8, 31);
		unsigned win = prev | (n8 << 8);          /* N flags of bases 8w-8 .. 8w+7 */
		unsigned s = win | (win << 1);
		s |= s << 2;
		s |= s << 4;
		s |= win << 8;                            /* bit b: an N among bases b-8 .. b */
		unsigned bad = (s >> 8) & 0xFFu;
		if (w == 0)
			bad = 0xFFu;
		if (w < nw)
			inval[w] = (uint8_t) bad;
	}
}

/* ---- K1-K3 of align() (assembler.c:92-118): flag every overlap that shares a valid 8-mer between the reads ----
 * HASN: some base is N, so k-mer validity comes from the inval bitmaps (warp-uniform choice, made per pair). */
template <int ML, bool HASN>
__device__ __forceinline__ void seed_candidates(WarpSmem<ML> &ws, int F, int R, int mo, int nbits, int lane) {
	using WS = WarpSmem<ML>;
	constexpr unsigned IDXMASK = WS::NB - 1;
	constexpr unsigned PLIM = 1u << WS::PBITS;       /* entry ^ tag below this: the tags agree and the value is a position */
	uint16_t *bt16 = reinterpret_cast<uint16_t *>(ws.btab);
	unsigned ovf = 0;
	/* K1: forward 8-mers into the bucket table.  A lane claims its slot with one shared-memory atomic on the bucket's
	 * 4-bit counter; the order of the entries inside a bucket is whatever the atomics made it, which is why the probe
	 * below takes the two LOWEST matching positions (= the reference's "first two", assembler.c:93-100). */
	for (int p = 8 + lane; p - lane < F; p += 32) {
		bool live = p < F;
		if (HASN)
			live = live && ((ws.inval_f[p >> 3] >> (p & 7)) & 1u) == 0;
		const unsigned code = ws.code_f[p];
		const unsigned idx = code & IDXMASK;
		if (live) {
			const unsigned sh = 4u * (idx & 7u);
			const unsigned slot = (atomicAdd(&ws.bcnt[idx >> 3], 1u << sh) >> sh) & 15u;
			if (slot < 4u)
				bt16[idx * 4 + slot] = (uint16_t) (((code >> WS::IDXBITS) << WS::PBITS) | (unsigned) p);
			else
				ovf = 1u;      /* a 5th entry; counters past this point may carry into their neighbours, nobody reads them */
		}
	}
	const bool overflow = __any_sync(FULL, ovf != 0);
	__syncwarp();
	const int cbase = F - mo;                  /* overlap - mo = cbase - p + e */
	if (!overflow) {
		/* K2: reverse 8-mers probe one bucket each (assembler.c:104-110); byte flags = BIT_LIST_SET */
#pragma unroll 2
		for (int e = 8 + lane; e - lane < R; e += 32) {
			bool live = e < R;
			if (HASN)
				live = live && ((ws.inval_r[e >> 3] >> (e & 7)) & 1u) == 0;
			const unsigned code = ws.code_r[e];
			const unsigned tagsh = (code >> WS::IDXBITS) << WS::PBITS;
			const uint2 b = *reinterpret_cast<const uint2 *>(&ws.btab[code & IDXMASK]);
			/* entry ^ tag is the stored position iff the tags agree, and at least 2^PBITS otherwise (an empty entry is 0xFFFF);
			 * the two lowest of the four 16-bit values, two at a time (VIMNMX.U16x2) */
			const unsigned t2 = tagsh * 0x10001u;
			const unsigned y0 = b.x ^ t2, y1 = b.y ^ t2;
			unsigned lo2, hi2;
			asm("min.u16x2 %0, %1, %2;" : "=r"(lo2) : "r"(y0), "r"(y1));
			asm("max.u16x2 %0, %1, %2;" : "=r"(hi2) : "r"(y0), "r"(y1));
			const unsigned l0 = lo2 & 0xFFFFu, l1 = lo2 >> 16, h0 = hi2 & 0xFFFFu, h1 = hi2 >> 16;
			const unsigned m1 = min(l0, l1), m2 = min(max(l0, l1), min(h0, h1));
			const unsigned c = live ? (unsigned) (cbase + e) : 0u;   /* dead lanes: index underflows, no flag */
			const unsigned i1 = c - m1, i2 = c - m2;
			if (m1 < PLIM && i1 < (unsigned) nbits)
				ws.cflag[i1] = 1;
			if (m2 < PLIM && i2 < (unsigned) nbits)
				ws.cflag[i2] = 1;
		}
		__syncwarp();
		/* K3: clear (assembler.c:113-116); an empty entry is 0xFFFF */
		const uint4 z = make_uint4(0, 0, 0, 0), ones = make_uint4(~0u, ~0u, ~0u, ~0u);
		uint4 *t4 = reinterpret_cast<uint4 *>(ws.btab);
#pragma unroll
		for (int k = 0; k < WS::NB * 8 / 16 / 32; k++)
			t4[k * 32 + lane] = ones;
		uint4 *c4 = reinterpret_cast<uint4 *>(ws.bcnt);
#pragma unroll
		for (int k = 0; k < (WS::NB / 8 * 4 + 511) / 512; k++)
			if (k * 32 + lane < WS::NB / 8 / 4)
				c4[k * 32 + lane] = z;
		return;
	}
	for (int k = lane; k < WS::NB / 8; k += 32)
		ws.bcnt[k] = 0;
	/* Exact open-addressing path for pairs whose k-mers crowd a bucket (low-complexity reads): every
	 * (code, position) is stored; the probe walks the whole chain and keeps the two lowest positions. */
	constexpr unsigned SMASK = WS::SLOTS - 1;
	uint32_t *slots = reinterpret_cast<uint32_t *>(ws.btab);
	for (int k = lane; k < WS::SLOTS; k += 32)
		slots[k] = 0;
	__syncwarp();
	for (int base = 8; base < F; base += 32) {
		const int p = base + lane;
		bool live = p < F;
		if (HASN)
			live = live && ((ws.inval_f[p >> 3] >> (p & 7)) & 1u) == 0;
		const unsigned code = ws.code_f[p];
		const unsigned entry = (code << 16) | (unsigned) p;
		unsigned slot = hash_slot(code, SMASK);
		bool pending = live;
		while (__any_sync(FULL, pending)) {
			bool tryw = false;
			if (pending) {
				tryw = slots[slot] == 0u;
				if (tryw)
					slots[slot] = entry;
			}
			__syncwarp();
			if (pending) {
				if (tryw && slots[slot] == entry)
					pending = false;
				else
					slot = (slot + 1) & SMASK;
			}
			__syncwarp();
		}
	}
	__syncwarp();
	for (int base = 8; base < R; base += 32) {
		const int e = base + lane;
		bool live = e < R;
		if (HASN)
			live = live && ((ws.inval_r[e >> 3] >> (e & 7)) & 1u) == 0;
		if (live) {
			const unsigned code = ws.code_r[e];
			unsigned slot = hash_slot(code, SMASK);
			unsigned m1 = 0x7FFFu, m2 = 0x7FFFu;
			for (;;) {
				const unsigned ent = slots[slot];
				if (ent == 0u)
					break;
				if ((ent >> 16) == code) {
					const unsigned p = ent & 0xFFFFu;
					if (p < m1) {
						m2 = m1;
						m1 = p;
					} else if (p < m2) {
						m2 = p;
					}
				}
				slot = (slot + 1) & SMASK;
			}
			const unsigned c = (unsigned) (cbase + e);
			const unsigned i1 = c - m1, i2 = c - m2;
			if (i1 < (unsigned) nbits)
				ws.cflag[i1] = 1;
			if (i2 < (unsigned) nbits)
				ws.cflag[i2] = 1;
		}
	}
	__syncwarp();
	for (int k = lane; k < WS::SLOTS; k += 32)
		slots[k] = ~0u;                /* back to the bucket table's "empty" */
}

/* ---- K6 of align() (assembler.c:158-244), 8 output bases per lane ----
 * Output position idx takes forward base fo+idx while idx < fend and reverse (template) base idx-df from
 * idx >= dfp on; where both apply the bases are merged (assembler.c:181-228).
 * GENERAL = false is the common case: no B-cliff masking, no degenerate bases, per-base p not requested. */
struct ReconArgs {
	const uint32_t *fnt32, *rnt32;
	const int8_t *fq, *rq;
	int fo, df, dfp, fend, seq_len, unmasked_f, lead_r, out_cap;
	uint8_t *out_nt;
	double *out_p;
	uint16_t *out_code;            /* instead of out_p: the per-base log p as its index into recon[2][48][48] (what the object layer ships to the host) */
	const uint16_t *qoff;          /* [0..255]: clamp(q)*48*8, [256..511]: clamp(q)*8, indexed by the raw quality byte */
};

constexpr unsigned ZERO_OFF = 2 * PB_NQM * PB_NQM * 8;   /* byte offset of the 0.0 that follows recon[2][48][48] in shared memory */

template <bool GENERAL>
__device__ __forceinline__ double recon_words(const ReconArgs &ra, const double *__restrict__ s_recon, int &mism, int &degen, int lane) {
	double qsum = 0.0;
	const int nwords = (ra.seq_len + 7) >> 3;
	for (int k = lane; k - lane < nwords; k += 32) {
		if (k < nwords) {
			const int idx0 = 8 * k;
			const int nV = min(ra.seq_len - idx0, 8);                 /* positions of this word inside the sequence */
			const int nF = min(max(ra.fend - idx0, 0), 8);            /* leading positions that use the forward read */
			const int r0 = min(max(ra.dfp - idx0, 0), 8);             /* positions < r0 do not use the reverse read */
			const unsigned maskV = nibmask(nV), maskF = nibmask(nF) & maskV, maskR = maskV & ~nibmask(r0);
			const unsigned fw = nibwin(ra.fnt32, ra.fo + idx0) & maskF;
			const unsigned rw = nibwin(ra.rnt32, idx0 - ra.df) & maskR;
			const unsigned both = maskF & maskR;
			const unsigned andw = fw & rw;
			const unsigned nzb = nz_nib(andw);                     /* match flags, meaningful where both */
			unsigned missb = both & NIB1 & ~nzb;                   /* mismatching overlap positions */
			/* match -> intersection; mismatch -> forward unless the reverse quality is strictly higher (fixed below) */
			unsigned nt = (fw & ~maskR) | (rw & ~maskF) | andw | (fw & (missb * 15u));
			const int8_t *fqp = ra.fq + ra.fo + idx0, *rqp = ra.rq + idx0 - ra.df;
			mism += __popc(missb);
			while (missb) {                                        /* assembler.c:215-219 */
				const int t = (__ffs(missb) - 1) >> 2;
				missb &= missb - 1;
				if (fqp[t] < rqp[t])
					nt = (nt & ~(15u << (4 * t))) | (rw & (15u << (4 * t)));
			}
			if (GENERAL)
				degen += __popc(~onehot_nib(nt) & maskV & NIB1);
			if (ra.out_nt && idx0 < ra.out_cap)
				reinterpret_cast<uint32_t *>(ra.out_nt)[k] = nt;
			/* per-base posterior: recon[match][a][b], a/b = clamped PHRED or 47 when that read is absent/masked.
			 * qoff maps a raw quality byte to the BYTE offset of its clamped row (x 48*8) / column (x 8). */
			const unsigned char *fqu = reinterpret_cast<const unsigned char *>(fqp), *rqu = reinterpret_cast<const unsigned char *>(rqp);
			const char *tab = reinterpret_cast<const char *>(s_recon);
#pragma unroll
			for (int t = 0; t < 8; t++) {
				unsigned oa = ra.qoff[fqu[t]], ob = ra.qoff[256 + rqu[t]];
				if (t >= nF)
					oa = PB_NQ * PB_NQM * 8;
				if (t < r0)
					ob = PB_NQ * 8;
				if (GENERAL) {                                     /* assembler.c:194-210 */
					const bool inboth = t < nF && t >= r0;
					if (inboth && ra.fo + idx0 + t >= ra.unmasked_f)
						oa = PB_NQ * PB_NQM * 8;
					if (inboth && idx0 + t - ra.df < ra.lead_r)
						ob = PB_NQ * 8;
				}
				const unsigned om = (nzb & (1u << (4 * t))) ? (unsigned) (PB_NQM * PB_NQM * 8) : 0u;
				unsigned off = oa + ob + om;
				if (t >= nV)
					off = ZERO_OFF;                                /* a 0.0 kept right after the table */
				const double p = *reinterpret_cast<const double *>(tab + off);
				qsum += p;
				if (GENERAL && ra.out_p && t < nV && idx0 + t < ra.out_cap)
					ra.out_p[idx0 + t] = p;
				if (GENERAL && ra.out_code && t < nV && idx0 + t < ra.out_cap)
					ra.out_code[idx0 + t] = (uint16_t) (off >> 3);
			}
		}
	}
	return qsum;
}

/* FULLF = false is the lean instantiation: no primers (before or after assembly) and no log()-based scorers
 * (ea_util, stitch); the host picks it whenever the configuration allows, which keeps the common kernel small. */
/* hang.c:39-72: the overhang trimmer, applied to the staged record in place.  The forward read keeps a prefix; the
 * reverse read keeps a prefix in READ order, i.e. a suffix of its template-order copy, which is moved to the front.
 * Returns false when the pair is dropped (sequence not found and hang_skip unset). */
template <int ML>
__device__ bool trim_overhangs(uint8_t *rec, int F0, int R0, int &F, int &R, const pb_device_params *__restrict__ prm,
                               const double *__restrict__ s_score, const uint16_t *__restrict__ s_qoff, double *se, int lane) {
	const int fwb = ((F0 + 7) / 8) * 4, rwb = ((R0 + 7) / 8) * 4;
	uint32_t *fnt32 = reinterpret_cast<uint32_t *>(rec), *rnt32 = reinterpret_cast<uint32_t *>(rec + fwb);
	int8_t *fq = reinterpret_cast<int8_t *>(rec + fwb + rwb), *rq = fq + ((F0 + 3) / 4) * 4;
	F = F0;
	R = R0;
	if (prm->hang_forward_length > 0) {
		/* index i of the scan is read position F-1-i: the "template order" form of primer_offset */
		const int off = primer_offset<true>(fnt32, fq, F0, prm->hang_forward, prm->hang_forward_length, prm->hang_threshold, 0.0, s_score, s_qoff, se, lane);
		if (off == 0) {
			if (!prm->hang_skip)
				return false;
		} else {
			F = F0 - (off - 1);
		}
	}
	if (prm->hang_reverse_length > 0) {
		/* the reverse read is stored in template order: scan index i (read position R-1-i) is element i */
		const int off = primer_offset<false>(rnt32, rq, R0, prm->hang_reverse, prm->hang_reverse_length, prm->hang_threshold, 0.0, s_score, s_qoff, se, lane);
		if (off == 0) {
			if (!prm->hang_skip)
				return false;
		} else {
			R = R0 - (off - 1);
		}
	}
	__syncwarp();
	if (F < F0) {        /* bases past the new end must read as padding */
		const int w = F >> 3;
		if (lane == 0) {
			fnt32[w] &= nibmask(F & 7);
			if (w + 1 < fwb / 4)
				fnt32[w + 1] = 0;
		}
	}
	if (R < R0) {
		const int d = R0 - R, nw = (R + 7) >> 3, nq = (R + 3) >> 2;
		uint32_t nt_new[(ML + 7) / 8 / 32 + 1], q_new[(ML + 3) / 4 / 32 + 1];
		const uint32_t *rq32 = reinterpret_cast<const uint32_t *>(rq);
#pragma unroll
		for (int k = 0; k < (ML + 7) / 8 / 32 + 1; k++) {
			const int w = k * 32 + lane;
			nt_new[k] = w < nw ? (nibwin(rnt32, 8 * w + d) & nibmask(min(R - 8 * w, 8))) : 0u;
		}
#pragma unroll
		for (int k = 0; k < (ML + 3) / 4 / 32 + 1; k++) {
			const int w = k * 32 + lane;
			unsigned v = 0;
			if (w < nq) {
				const int b = 4 * w + d;
				v = __funnelshift_r(rq32[b >> 2], rq32[(b >> 2) + 1], (b & 3) * 8);
				const int keep = min(R - 4 * w, 4);
				if (keep < 4)
					v &= (1u << (8 * keep)) - 1u;
			}
			q_new[k] = v;
		}
		__syncwarp();
#pragma unroll
		for (int k = 0; k < (ML + 7) / 8 / 32 + 1; k++) {
			const int w = k * 32 + lane;
			if (w < rwb / 4)
				rnt32[w] = nt_new[k];
		}
		uint32_t *rq32w = reinterpret_cast<uint32_t *>(rq);
#pragma unroll
		for (int k = 0; k < (ML + 3) / 4 / 32 + 1; k++) {
			const int w = k * 32 + lane;
			if (w < ((R0 + 3) >> 2))
				rq32w[w] = q_new[k];
		}
	}
	__syncwarp();
	return true;
}

template <int ML, bool FULLF>
__device__ void process_pair(WarpSmem<ML> &ws, uint8_t *rec, int F, int R,
                             const pb_device_params *__restrict__ prm,
                             const double *__restrict__ s_recon, const double *__restrict__ s_over,
                             const double *__restrict__ s_score, const double *__restrict__ s_score_err,
                             const uint16_t *__restrict__ s_qoff, const uint8_t *__restrict__ s_primer, uint8_t *scratch,
                             pb_pair_result &res, uint8_t *out_nt, double *out_p, uint16_t *out_code, int out_cap, int lane) {
	using WS = WarpSmem<ML>;
	PairView v;
	const int fwb = ((F + 7) / 8) * 4, rwb = ((R + 7) / 8) * 4;
	v.fnt = rec;
	v.rnt = rec + fwb;
	v.fnt32 = reinterpret_cast<const uint32_t *>(v.fnt);
	v.rnt32 = reinterpret_cast<const uint32_t *>(v.rnt);
	v.fq = (const int8_t *) (rec + fwb + rwb);
	v.rq = v.fq + ((F + 3) / 4) * 4;
	/* The primer / overhang scans keep two doubles per read position in the bucket table's memory (it is idle until the
	 * k-mer join); BucketRestore puts the table's "empty" pattern back on every way out of this function. */
	static_assert(sizeof(ws.btab) >= (size_t) ML * 16 + (size_t) ML * 4, "the bucket table must hold two doubles per read position and the primer masks");
	double *const se = reinterpret_cast<double *>(ws.btab);
	struct BucketRestore {
		WarpSmem<ML> &w;
		bool on;
		int lane;
		__device__ void now() {
			if (on) {
				__syncwarp();
				uint4 *t4 = reinterpret_cast<uint4 *>(w.btab);
				const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
				for (int k = 0; k < WarpSmem<ML>::NB * 8 / 16 / 32; k++)
					t4[k * 32 + lane] = ones;
				__syncwarp();
				on = false;
			}
		}
		__device__ ~BucketRestore() { now(); }      /* the early returns (NOFP, NORP, BADR, NOALGN, dropped by the trimmer) */
	};
	BucketRestore restore{ ws, FULLF && (prm->hang_forward_length > 0 || prm->hang_reverse_length > 0 || (prm->post_primers == 0 && (prm->forward_primer_length > 0 || prm->reverse_primer_length > 0))), lane };
	if (FULLF && (prm->hang_forward_length > 0 || prm->hang_reverse_length > 0)) {
		int Ft, Rt;
		if (!trim_overhangs<ML>(rec, F, R, Ft, Rt, prm, s_score, s_qoff, se, lane)) {
			res.status = PB_PAIR_SKIP;         /* the reader drops the pair: the assembler never sees it */
			res.slow = 0;
			res.overlap = res.seq_len = res.mismatches = res.degenerates = res.examined = 0;
			res.fwd_offset = res.rev_offset = 0;
			res.quality = 0.0;
			res.est_prob = 0.0;
			return;
		}
		F = Ft;
		R = Rt;
	}
	v.F = F;
	v.R = R;

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
	const bool post = FULLF && prm->post_primers != 0;   /* assembler.c:262,285-288: primers are located after assembly instead */
	const bool staged = FULLF && (post || prm->need_stage != 0);   /* the assembled sequence goes through this warp's scratch */
	/* assembler.c:262-284 (primers before assembly) */
	if (post) {
		fo = 0;
		ro = 0;
	} else {
	if (FULLF && prm->forward_primer_length > 0) {
		fo = primer_offset<false>(v.fnt32, v.fq, F, s_primer, prm->forward_primer_length,
		                          prm->threshold, prm->primer_penalty, s_score, s_qoff, se, lane);
		if (fo == 0) {
			res.status = PB_PAIR_NOFP;
			return;
		}
		fo--;
	} else {
		fo = prm->forward_trim;
	}
	res.fwd_offset = (uint16_t) fo;
	if (FULLF && prm->reverse_primer_length > 0) {
		ro = primer_offset<true>(v.rnt32, v.rq, R, s_primer + PB_MAX_LEN + 2, prm->reverse_primer_length,
		                         prm->threshold, prm->primer_penalty, s_score, s_qoff, se, lane);
		if (ro == 0) {
			res.status = PB_PAIR_NORP;
			return;
		}
		ro--;
	} else {
		ro = prm->reverse_trim;
	}
	res.rev_offset = (uint16_t) ro;
	}
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

	restore.now();       /* the scans are over: the bucket table is a bucket table again */
	/* ---- k-mer codes of both reads ---- */
	unsigned flg = gen_codes<WS::NTW>(v.fnt32, F, ws.code_f, lane) | gen_codes<WS::NTW>(v.rnt32, R, ws.code_r, lane);
	flg = __reduce_or_sync(FULL, flg);
	const bool anyN = (flg & 1u) != 0;
	const bool anyDeg = (flg & 2u) != 0;       /* some base is not A/C/G/T: only then can a merged base be degenerate */
	if (anyN) {
		gen_invalid<WS::NTW>(v.fnt32, F, ws.inval_f, lane);
		gen_invalid<WS::NTW>(v.rnt32, R, ws.inval_r, lane);
	}
	__syncwarp();

	/* ---- K1-K3: seed candidate overlaps (assembler.c:92-118) ---- */
	if (anyN)
		seed_candidates<ML, true>(ws, F, R, mo, nbits, lane);
	else
		seed_candidates<ML, false>(ws, F, R, mo, nbits, lane);

	/* ---- K4/K5: sweep the candidates in increasing overlap (assembler.c:118-143) ---- */
	const double qual_nn = prm->qual_nn;
	double best = qual_nn * (double) (unsigned long long) (F + R);   /* assembler.c:60 */
	int bestov = -1;
	int examined = 0;
	const int nchunks = (nbits + 15) >> 4;     /* 16 flag bytes per chunk; nbits < 900 -> at most 57 chunks */
	const int algo = prm->algo;
	unsigned cm[2];
#pragma unroll
	for (int h = 0; h < 2; h++) {
		unsigned nzw = 0;
		if (h * 32 < nchunks && h * 32 + lane < nchunks) {
			const uint4 f = reinterpret_cast<const uint4 *>(ws.cflag)[h * 32 + lane];
			nzw = f.x | f.y | f.z | f.w;
		}
		cm[h] = (h * 32 < nchunks) ? __ballot_sync(FULL, nzw != 0) : 0u;
	}
	const bool none = (cm[0] | cm[1]) == 0;    /* ALL_BITS_IF_NONE, assembler.c:118 */
	if (none) {
		cm[0] = nchunks >= 32 ? FULL : ((1u << nchunks) - 1u);
		cm[1] = nchunks > 32 ? ((1u << (nchunks - 32)) - 1u) : 0u;
	}
#pragma unroll
	for (int h = 0; h < 2; h++) {
		unsigned chunks = cm[h];
		while (chunks) {
			const int ch = h * 32 + __ffs(chunks) - 1;
			chunks &= chunks - 1;
			/* candidate bits of this chunk: lane t < 16 looks at flag byte t */
			const int idx0 = ch * 16;
			bool set = false;
			if (lane < 16 && idx0 + lane < nbits)
				set = none || ws.cflag[idx0 + lane] != 0;
			unsigned cand = __ballot_sync(FULL, set);
			__syncwarp();
			if (!none && lane == 0)
				reinterpret_cast<uint4 *>(ws.cflag)[ch] = make_uint4(0, 0, 0, 0);
			while (cand) {
				const int ov = idx0 + __ffs(cand) - 1 + mo;
				cand &= cand - 1;
				/* overlap_probability for this candidate.  findex = F-ov+i in [0,F), template index i in [0,R) */
				const int i0 = max(0, ov - F), i1 = min(ov, R);
				double prob;
				if (algo != PB_PEAR && algo != PB_RDP_MLE) {      /* the count-based scorers */
					unsigned packed = 0;
					const int nw = (i1 + 7) >> 3;
					for (int k0 = 0; k0 < nw; k0 += 32) {
						const int k = k0 + lane;
						if (k < nw) {
							const unsigned M = (nibmask(min(max(i1 - 8 * k, 0), 8)) & ~nibmask(min(max(i0 - 8 * k, 0), 8))) & NIB1;
							const unsigned r = v.rnt32[k];
							const unsigned f = nibwin(v.fnt32, F - ov + 8 * k);
							const unsigned nzb = nz_nib(f & r);
							const unsigned unk = anyN ? ((n_nib(f) | n_nib(r)) & M) : 0u;
							const unsigned mt = nzb & ~unk & M, mm = ~nzb & ~unk & M;
							packed += (unsigned) __popc(mt) | ((unsigned) __popc(mm) << 10) | ((unsigned) __popc(unk) << 20);
						}
					}
					packed = __reduce_add_sync(FULL, packed);
					const int matches = packed & 1023, mism = (packed >> 10) & 1023, unk = packed >> 20;
					if (algo == PB_SIMPLE_BAYES || algo == PB_UPARSE) {
						/* algo_simple_bayes.c:61-65 / algo_uparse.c:61-65: size_t arithmetic inside the parenthesis
						 * (sb_pmatch / sb_pmismatch hold the selected algorithm's two constants) */
						const unsigned long long nn_count = (ov >= F && ov >= R)
							? (unsigned long long) unk
							: (unsigned long long) ((long long) F + R - 2 * (long long) ov + unk);
						prob = qual_nn * (double) nn_count + (double) matches * prm->sb_pmatch;
						prob = prob + (double) mism * prm->sb_pmismatch;
					} else if (algo == PB_FLASH) {
						/* algo_flash.c:59: integer division inside log() */
						const int real = matches + mism + unk, bad = mism + unk;
						prob = (real == 0) ? -2.0 : ((bad == real) ? 0.0 : -CUDART_INF);
					} else if (FULLF && algo == PB_EA_UTIL) {
						/* algo_ea_util.c:55: N counts as a mismatch; real_overlap == 0 divides by zero as the reference does */
						const double bad = (double) (mism + unk);
						prob = log((bad * bad + 1.0) / (double) (unsigned long long) (matches + mism + unk));
					} else if (FULLF) {
						/* algo_stitch.c:55: the score is a size_t, a net-negative score wraps */
						const unsigned long long sc = (unsigned long long) (long long) (matches - mism);
						prob = log((double) sc / (double) (unsigned long long) (F + R));
					} else {
						prob = -CUDART_INF;      /* not reachable: the host never pairs these scorers with the lean kernel */
					}
				} else {
					double acc = 0.0;
					for (int i = i0 + lane; i < i1; i += 32) {
						const int fi = F - ov + i;
						const unsigned f = nib(v.fnt, fi), r = nib(v.rnt, i);
						const int qa = clampq(v.fq[fi]);
						if (algo == PB_PEAR) {
							/* algo_pear.c:52,54 index the FORWARD qualities with rindex = R-1-i; past the
							 * end of the forward read that is defined as quality 0.
							 * pear's overlap terms are the very matrices its match_probability uses, so they are
							 * read from the reconstruction table (row stride 48) and need no table of their own. */
							const int ri = R - 1 - i;
							const int qb = (ri < F) ? clampq(v.fq[ri]) : 0;
							if (f == 15u || r == 15u)
								acc -= prm->pear_random_base;
							else
								acc += s_recon[(((f & r) ? 1 : 0) * PB_NQM + qa) * PB_NQM + qb];
						} else {
							const int qb = clampq(v.rq[i]);
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
	if (v.fq[F - 1] == 2 || v.rq[0] == 2) {       /* warp-uniform: most pairs have no '#' tail */
		for (int base = 0; base < F; base += 32) {
			const int i = F - 1 - base - lane;
			const unsigned notb = __ballot_sync(FULL, !(i >= 0 && v.fq[max(i, 0)] == 2));
			if (notb) {
				unmasked_f = F - base - (__ffs(notb) - 1);
				break;
			}
			unmasked_f = max(F - base - 32, 0);
		}
		for (int base = 0; base < R; base += 32) {
			const int j = base + lane;      /* template order: reverse[R-1-j] */
			const unsigned notb = __ballot_sync(FULL, !(j < R && v.rq[min(j, R - 1)] == 2));
			if (notb) {
				lead_r = base + (__ffs(notb) - 1);
				break;
			}
			lead_r = min(base + 32, R);
		}
	}
	const bool cliff = unmasked_f < F || lead_r > 0;
	/* the per-word work, specialised on the (warp-uniform, rare) B-cliff / per-base-p / degenerate cases */
	double qsum;
	int mism = 0, degen = 0;
	{
		ReconArgs ra;
		ra.fnt32 = v.fnt32; ra.rnt32 = v.rnt32; ra.fq = v.fq; ra.rq = v.rq;
		ra.fo = fo; ra.df = df; ra.dfp = dfp; ra.fend = dfp + nover; ra.seq_len = seq_len;
		ra.unmasked_f = unmasked_f; ra.lead_r = lead_r; ra.out_nt = out_nt; ra.out_p = out_p; ra.out_code = out_code; ra.out_cap = out_cap; ra.qoff = s_qoff;
		if (staged) {      /* the whole assembled sequence goes to this warp's scratch first */
			ra.out_p = reinterpret_cast<double *>(scratch);
			ra.out_code = nullptr;         /* the host asks for codes only when nothing is staged (pb_device.cu) */
			ra.out_nt = scratch + 912 * 8;
			ra.out_cap = 912;
		}
		if (cliff || anyDeg || ra.out_p != nullptr || ra.out_code != nullptr)
			qsum = recon_words<true>(ra, s_recon, mism, degen, lane);
		else
			qsum = recon_words<false>(ra, s_recon, mism, degen, lane);
	}
	qsum = warp_sum(qsum);
	mism = __reduce_add_sync(FULL, mism);
	degen = __reduce_add_sync(FULL, degen);
	res.quality = qsum / (double) len;               /* assembler.c:244: divides by len, not seq_len */
	res.overlap = (uint16_t) bestov;
	res.est_prob = best;
	res.seq_len = (uint16_t) seq_len;
	res.mismatches = (uint16_t) mism;
	res.degenerates = (uint16_t) degen;
	int min_phred = 127;
	if (FULLF && staged) {                           /* assembler.c:300-333 */
		__syncwarp();
		const double *sp = reinterpret_cast<const double *>(scratch);
		const uint32_t *snt32 = reinterpret_cast<const uint32_t *>(scratch + 912 * 8);
		int pfo = 0, pro = 0;
		if (!post) {
			/* staged only because a filter reads the per-base log p: nothing is stripped */
		} else if (prm->forward_primer_length > 0) {
			pfo = primer_offset_result(snt32, sp, seq_len, false, s_primer, prm->forward_primer_length, prm->threshold, prm->primer_penalty, lane);
			if (pfo == 0) {
				res.status = PB_PAIR_NOFP;
				return;
			}
			pfo--;
		} else {
			pfo = prm->forward_trim;
		}
		if (post)
			res.fwd_offset = (uint16_t) pfo;
		if (!post) {
		} else if (prm->reverse_primer_length > 0) {
			pro = primer_offset_result(snt32, sp, seq_len, true, s_primer + PB_MAX_LEN + 2, prm->reverse_primer_length, prm->threshold, prm->primer_penalty, lane);
			if (pro == 0) {
				res.status = PB_PAIR_NORP;
				return;
			}
			pro--;
		} else {
			pro = prm->reverse_trim;
		}
		if (post)
			res.rev_offset = (uint16_t) pro;
		if (post && seq_len <= pfo + pro) {
			res.status = PB_PAIR_NOFP;               /* sic: assembler.c:324-328 */
			return;
		}
		const int newlen = seq_len - pfo - pro;      /* quality, degenerates, mismatches keep their pre-strip values */
		res.seq_len = (uint16_t) newlen;
		if (out_nt) {
			const int nw = (newlen + 7) >> 3;
			for (int k = lane; k < nw; k += 32)
				if (8 * k < out_cap)
					reinterpret_cast<uint32_t *>(out_nt)[k] = nibwin(snt32, pfo + 8 * k) & nibmask(min(newlen - 8 * k, 8));
		}
		if (out_p) {
			for (int k = lane; k < newlen; k += 32)
				if (k < out_cap)
					out_p[k] = sp[pfo + k];
		}
		if (prm->need_stage) {                       /* plugin_min_phred.c:15-20 over what is emitted */
			int mp = 127;
			for (int k = lane; k < newlen; k += 32) {
				/* nt.c:126-150 */
				const double p = sp[pfo + k];
				int lower = 0, upper = PB_PHREDMAX, ph = -1;
				if (p <= s_score[0])
					ph = 1;
				while (ph < 0 && lower < upper) {
					const int mid = lower + (upper - lower) / 2;
					const double sc = s_score[mid];
					if (sc == p)
						ph = mid;
					else if (mid == lower)
						ph = lower;
					else if (sc > p)
						upper = mid;
					else
						lower = mid + 1;
				}
				if (ph < 0)
					ph = lower;
				mp = min(mp, ph);
			}
			min_phred = __reduce_min_sync(FULL, mp);
		}
		__syncwarp();                                /* scratch is reused by this warp's next pair */
	}
	if (res.quality < prm->threshold) {              /* assembler.c:334-338 */
		res.status = PB_PAIR_LOWQ;
		return;
	}
	/* module_checkseq (module.c:124-137): the first failing check rejects the pair */
	const int nf = prm->nfilters;
	for (int k = 0; k < nf; k++) {
		const int kind = prm->filters[k].kind, iv = prm->filters[k].ivalue;
		bool pass = true;
		if (kind == PB_FILTER_NO_N)
			pass = res.degenerates == 0;
		else if (kind == PB_FILTER_SHORT)
			pass = (int) res.seq_len >= iv;
		else if (kind == PB_FILTER_LONG)
			pass = (int) res.seq_len <= iv;
		else if (kind == PB_FILTER_MIN_OVERLAPBITS)
			pass = prm->filters[k].dvalue * 0.693147180559945309417232121458 <= res.est_prob;     /* bits -> nats, M_LN2 */
		else if (kind == PB_FILTER_MISS_THE_POINT)
			pass = (int) res.mismatches <= iv;
		else if (kind == PB_FILTER_MIN_PHRED)
			pass = min_phred >= iv;
		else if (kind == PB_FILTER_PEAR_TEST)
			pass = pear_test_pass(prm->pear_cdf, prm->filters[k].dvalue, prm->filters[k].dvalue2, prm->filters[k].dvalue3,
			                      res.overlap, res.mismatches, F, R);
		if (!pass) {
			res.status = (uint8_t) (PB_PAIR_FILTERED + k);
			return;
		}
	}
}

template <int ML, bool OVER, int WARPS_PER_BLOCK, bool FULLF>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 1)
assemble_kernel(const pb_device_params *__restrict__ prm, int n,
                const uint8_t *__restrict__ reads, const pb_pair_meta *__restrict__ meta,
                pb_pair_result *__restrict__ results, uint8_t *__restrict__ seq_nt, double *__restrict__ seq_p,
                long long seq_stride, unsigned long long *__restrict__ counters, uint8_t *__restrict__ scratch_all,
                const int *__restrict__ list, const int *__restrict__ list_n, uint16_t *__restrict__ seq_code) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	using WS = WarpSmem<ML>;
	/* list mode: assemble pairs list[0 .. *list_n) -- the ones the lane-per-pair kernel (pb_lanes.cuh) deferred */
	if (list) {
		n = *list_n;
		if ((long long) blockIdx.x * WARPS_PER_BLOCK >= n)      /* usually a few hundred pairs: most CTAs have nothing to do */
			return;
	}
	/* block-level: LUTs + counters, then the per-warp areas */
	constexpr int OVER_N = OVER ? 2 * PB_NQ * PB_NQ : 0;
	double *s_recon = reinterpret_cast<double *>(smem_raw);
	double *s_over = s_recon + 2 * PB_NQM * PB_NQM + 2;        /* [2*48*48] is the 0.0 recon_words() uses for padding positions */
	double *s_score = s_over + OVER_N;
	double *s_score_err = s_score + PB_NQM;
	unsigned *s_cnt = reinterpret_cast<unsigned *>(s_score_err + PB_NQM);
	uint16_t *s_qoff = reinterpret_cast<uint16_t *>(s_cnt + PB_NCOUNTERS);
	uint8_t *s_primer = reinterpret_cast<uint8_t *>(s_qoff + 512);    /* forward primer, then the reverse one */
	constexpr size_t LUT_BYTES = (2 * PB_NQM * PB_NQM + 2 + OVER_N + 2 * PB_NQM) * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned) + 512 * sizeof(uint16_t) + 2 * (PB_MAX_LEN + 2);
	constexpr size_t LUT_ALIGNED = (LUT_BYTES + 127) & ~(size_t) 127;
	WS *wsall = reinterpret_cast<WS *>(smem_raw + LUT_ALIGNED);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < 2 * PB_NQM * PB_NQM; i += blockDim.x)
		s_recon[i] = (&prm->recon[0][0][0])[i];
	if (tid < 2)
		s_recon[2 * PB_NQM * PB_NQM + tid] = 0.0;
	if (OVER)
		for (int i = tid; i < 2 * PB_NQ * PB_NQ; i += blockDim.x)
			s_over[i] = (&prm->over[0][0][0])[i];
	for (int i = tid; i < PB_NQM; i += blockDim.x) {
		s_score[i] = prm->score[i];
		s_score_err[i] = prm->score_err[i];
	}
	for (int i = tid; i < PB_NCOUNTERS; i += blockDim.x)
		s_cnt[i] = 0;
	for (int i = tid; i < 2 * (PB_MAX_LEN + 2); i += blockDim.x)
		s_primer[i] = i < PB_MAX_LEN + 2 ? prm->forward_primer[i] : prm->reverse_primer[i - (PB_MAX_LEN + 2)];
	for (int i = tid; i < 256; i += blockDim.x) {
		const int q = clampq((int) (signed char) i);
		s_qoff[i] = (uint16_t) (q * PB_NQM * 8);
		s_qoff[256 + i] = (uint16_t) (q * 8);
	}
	WS &ws = wsall[warp];
	for (int k = lane; k < WS::NB; k += 32)
		ws.btab[k] = ~0ull;
	for (int k = lane; k < WS::NB / 8; k += 32)
		ws.bcnt[k] = 0;
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
	/* merged read rows: 4 bit per base, seq_stride bases per row */
	const long long nt_row = seq_stride / 2;
	uint8_t *scratch = scratch_all ? scratch_all + (size_t) wglobal * PB_SCRATCH_STRIDE : nullptr;

	auto issue = [&](int k, int stage) {
		if (lane == 0) {
			pb_pair_meta m = meta[list ? list[k] : k];
			ws.meta[stage] = m;
			unsigned bytes = m.flen == 0xFFFFu ? 0u : record_bytes(m.flen, m.rlen);     /* 0xFFFF: not a pair (FASTQ reader) */
			mbar_expect_tx(&ws.bar[stage], bytes);
			if (bytes)
				bulk_g2s(ws.stage[stage], reads + (size_t) m.off16 * 16, bytes, &ws.bar[stage]);
		}
	};

	int item = wglobal;
	if (item < n)
		issue(item, 0);
	for (int it = 0; item < n; it++, item += wstride) {
		const int stage = it & 1;
		const int next = item + wstride;
		if (next < n)
			issue(next, stage ^ 1);
		const int pair = list ? list[item] : item;
		mbar_wait(&ws.bar[stage], (it >> 1) & 1);
		const pb_pair_meta m = ws.meta[stage];
		union { pb_pair_result r; uint4 v[2]; } ru;
		pb_pair_result &res = ru.r;
		if (m.flen == 0xFFFFu) {       /* a record the FASTQ reader drops (fastq.c:176): no result, no counter */
			if (lane == 0) {
				ru.v[0] = make_uint4(0, 0, 0, 0);
				ru.v[1] = make_uint4(0, 0, 0, 0);
				res.status = PB_PAIR_SKIP;
				uint4 *dst = reinterpret_cast<uint4 *>(&results[pair]);
				dst[0] = ru.v[0];
				dst[1] = ru.v[1];
			}
			__syncwarp();
			continue;
		}
		uint8_t *o_nt = seq_nt ? seq_nt + (size_t) pair * nt_row : nullptr;
		double *o_p = seq_p ? seq_p + (size_t) pair * seq_stride : nullptr;
		uint16_t *o_c = seq_code ? seq_code + (size_t) pair * seq_stride : nullptr;
		process_pair<ML, FULLF>(ws, ws.stage[stage], m.flen, m.rlen, prm, s_recon, s_over, s_score, s_score_err, s_qoff, s_primer, scratch, res, o_nt, o_p, o_c, (int) seq_stride, lane);
		if (lane == 0) {
			uint4 *dst = reinterpret_cast<uint4 *>(&results[pair]);
			dst[0] = ru.v[0];
			dst[1] = ru.v[1];
			if (res.status != PB_PAIR_SKIP)
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
			case PB_PAIR_SKIP: break;
			default: atomicAdd(&s_cnt[PB_C_REJECTED + res.status - PB_PAIR_FILTERED], 1u); break;
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

/* ---- seeding on its own: K1-K3 of align() for every pair of the batch, one warp per pair ------------------------
 * First half of the two-kernel path for the common configurations (pb_lanes.cuh is the second half): only the packed
 * bases of a record are staged, the candidate overlaps come out as a bit mask per pair.
 *   seeds[pair][0..4]  bit i set <=> overlap minoverlap + i shares a valid 8-mer between the reads (BIT_LIST_SET)
 *   seeds[pair][5]     PB_SEED_GENERAL: leave this pair to the general kernel (a base that is not A/C/G/T, reads
 *                      outside 16..ML, more than 160 candidate overlaps, no seed at all);
 *                      PB_SEED_SKIP: not a pair (flen == 0xFFFF). */
constexpr unsigned PB_SEED_GENERAL = 1u, PB_SEED_SKIP = 2u;
/* words of a seeds record for reads up to ML nt: the mask, the flag word, the bin, padded to a multiple of four */
__host__ __device__ constexpr int seed_mask_words(int ML) { return (ML + 31) / 32; }
__host__ __device__ constexpr int seed_words(int ML) { return (seed_mask_words(ML) + 2 + 3) & ~3; }
/*   seeds[pair][MW+1]  (MW = mask words: flags sit at [MW]) bin of the pair: (lowest candidate overlap - minoverlap) / 16, or PB_SEED_BINS - 1 for the pairs that
 *                      carry a flag.  bin_order_kernel lists the pairs bin by bin, so that the 32 pairs a warp of the lane-per-
 *                      pair kernel takes have overlaps within 16 bases of each other and its loops end together. */
constexpr int PB_SEED_BINS = 21;      /* 20 x 16 overlaps (reads up to 320 nt) + the flagged pairs */

template <int ML, int WARPS_PER_BLOCK>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 1)
seed_kernel(const pb_device_params *__restrict__ prm, int n, const uint8_t *__restrict__ reads,
            const pb_pair_meta *__restrict__ meta, uint32_t *__restrict__ seeds, unsigned *__restrict__ bin_count) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	__shared__ unsigned s_bins[PB_SEED_BINS];
	using WS = WarpSmem<ML>;
	static_assert(ML <= 256, "bins and mask words are laid out for reads up to 256 nt");
	constexpr int MW = seed_mask_words(ML), SWORDS = seed_words(ML);
	WS *wsall = reinterpret_cast<WS *>(smem_raw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < PB_SEED_BINS)
		s_bins[tid] = 0;
	WS &ws = wsall[warp];
	for (int k = lane; k < WS::NB; k += 32)
		ws.btab[k] = ~0ull;
	for (int k = lane; k < WS::NB / 8; k += 32)
		ws.bcnt[k] = 0;
	for (int k = lane; k < WS::NFLAG; k += 32)
		ws.cflag[k] = 0;
	if (lane == 0) {
		for (int s = 0; s < NSTAGE; s++)
			mbar_init(&ws.bar[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const int mo = prm->minoverlap, cfg_maxov = prm->maxoverlap;
	const int wglobal = blockIdx.x * WARPS_PER_BLOCK + warp;
	const int wstride = gridDim.x * WARPS_PER_BLOCK;

	auto issue = [&](int pair, int stage) {
		if (lane == 0) {
			pb_pair_meta m = meta[pair];
			ws.meta[stage] = m;
			unsigned bytes = 0;
			if (m.flen != 0xFFFFu && m.flen <= ML && m.rlen <= ML)      /* only the packed bases: the qualities play no part in seeding */
				bytes = ((((unsigned) m.flen + 7) / 8) * 4 + (((unsigned) m.rlen + 7) / 8) * 4 + 15u) & ~15u;
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
		const int F = m.flen, R = m.rlen;
		unsigned word = 0, flags = 0, bin = PB_SEED_BINS - 1;
		if (F == 0xFFFF) {
			flags = PB_SEED_SKIP;
		} else if (F > ML || R > ML || F < 16 || R < 16 || mo >= min(F, R)) {
			flags = PB_SEED_GENERAL;
		} else {
			const int maxov = cfg_maxov == 0 ? min(F, R) : min(F + R - mo - 1, cfg_maxov);      /* assembler.c:59,78-84 */
			const int nbits = (mo <= maxov) ? (maxov - mo + 1) : 1;
			const uint32_t *fnt32 = reinterpret_cast<const uint32_t *>(ws.stage[stage]);
			const uint32_t *rnt32 = fnt32 + ((F + 7) >> 3);
			unsigned flg = gen_codes<WS::NTW>(fnt32, F, ws.code_f, lane) | gen_codes<WS::NTW>(rnt32, R, ws.code_r, lane);
			flg = __reduce_or_sync(FULL, flg);
			__syncwarp();
			if (flg != 0u || nbits > 32 * MW) {
				flags = PB_SEED_GENERAL;
			} else {
				seed_candidates<ML, false>(ws, F, R, mo, nbits, lane);
				__syncwarp();
				/* flag bytes -> mask words; lane w keeps word w */
				unsigned any = 0;
#pragma unroll
				for (int w = 0; w < MW; w++) {
					const unsigned b = __ballot_sync(FULL, ws.cflag[32 * w + lane] != 0);
					any |= b;
					if (lane == w)
						word = b;
				}
				/* the lowest candidate decides the bin */
				bin = __reduce_min_sync(FULL, word ? (unsigned) (32 * lane + __ffs(word) - 1) : 1023u) >> 4;
				__syncwarp();
				if (lane < 2 * MW)
					reinterpret_cast<uint4 *>(ws.cflag)[lane] = make_uint4(0, 0, 0, 0);
				if (any == 0u)
					flags = PB_SEED_GENERAL;       /* ALL_BITS_IF_NONE (assembler.c:118): every overlap is scored */
			}
		}
		if (flags)
			bin = PB_SEED_BINS - 1;
		if (lane == MW)
			word = flags;
		if (lane == MW + 1)
			word = bin;
		if (lane < SWORDS)
			seeds[(size_t) pair * SWORDS + lane] = lane < MW + 2 ? word : 0u;
		if (lane == 0)
			atomicAdd(&s_bins[bin], 1u);
		__syncwarp();      /* every lane is done with this stage before it is refilled */
	}
	__syncthreads();
	if (tid < PB_SEED_BINS && s_bins[tid])
		atomicAdd(&bin_count[tid], s_bins[tid]);
}

/* The pairs of a batch listed bin by bin (seeds[pair][6]); the order inside a bin is whatever the atomics make it, which no
 * result depends on.  bin_state: [0 .. BINS) pairs per bin (seed_kernel), [BINS .. 2 BINS) cursors, zero at launch. */
__global__ void __launch_bounds__(256)
bin_order_kernel(int n, const uint32_t *__restrict__ seeds, int swords, int bin_word, unsigned *__restrict__ bin_state, int *__restrict__ order,
                 const int *__restrict__ list, const int *__restrict__ list_n) {
	__shared__ unsigned s_hist[PB_SEED_BINS], s_base[PB_SEED_BINS];
	const int tid = threadIdx.x;
	if (list_n)
		n = *list_n;          /* the pairs of one length class, listed by index (class_list_kernel) */
	if ((int) (blockIdx.x * blockDim.x) >= n)
		return;
	if (tid < PB_SEED_BINS)
		s_hist[tid] = 0;
	__syncthreads();
	const int item = blockIdx.x * blockDim.x + tid;
	const int pair = item < n ? (list ? list[item] : item) : -1;
	unsigned bin = 0, rank = 0;
	if (pair >= 0) {
		bin = min(seeds[(size_t) pair * swords + bin_word], (unsigned) (PB_SEED_BINS - 1));
		rank = atomicAdd(&s_hist[bin], 1u);
	}
	__syncthreads();
	if (tid < PB_SEED_BINS) {
		unsigned first = 0;                       /* pairs in the bins before this one */
		for (int b = 0; b < tid; b++)
			first += bin_state[b];
		s_base[tid] = first + (s_hist[tid] ? atomicAdd(&bin_state[PB_SEED_BINS + tid], s_hist[tid]) : 0u);
	}
	__syncthreads();
	if (pair >= 0)
		order[s_base[bin] + rank] = pair;
}

/* Batches of mixed read lengths: the pairs listed by length class -- longest read of the pair <= 160, <= 256, <= 320 nt -- so that
 * every class runs the seeding sweep and the lane-per-pair kernel sized for it (the sweep's work grows with the square of the
 * length class, the lane kernel's shared memory per pair with the class).  lists[c * cap ..] = indices of class c, counts[c] their
 * number (zero at launch); pairs with a longer read go straight to the general kernel's list.  Order inside a list: whatever the
 * atomics make it; no result depends on it.  (Shared-memory ranks, one global atomic per class and block: with one per warp the
 * three hot counters cost 64 ns per thousand pairs.) */
constexpr int PB_LEN_CLASSES = 3;
__host__ __device__ constexpr int len_class_max(int c) { return c == 0 ? 160 : (c == 1 ? 256 : 320); }
__global__ void __launch_bounds__(256)
class_list_kernel(int n, const pb_pair_meta *__restrict__ meta, int *__restrict__ lists, size_t cap, int *__restrict__ counts,
                  int *__restrict__ general_list, int *__restrict__ general_count, unsigned long long *__restrict__ general_total) {
	__shared__ int s_n[PB_LEN_CLASSES + 1], s_base[PB_LEN_CLASSES + 1];
	const int tid = threadIdx.x;
	if (tid <= PB_LEN_CLASSES)
		s_n[tid] = 0;
	__syncthreads();
	const int pair = blockIdx.x * blockDim.x + tid;
	int cls = -1, rank = 0;
	if (pair < n) {
		const uint2 m = *reinterpret_cast<const uint2 *>(&meta[pair]);
		const int F = (int) (m.y & 0xFFFFu), R = (int) (m.y >> 16);
		const int longest = F == 0xFFFF ? 0 : max(F, R);        /* not a pair: any class reports it as such */
		cls = longest <= len_class_max(0) ? 0 : (longest <= len_class_max(1) ? 1 : (longest <= len_class_max(2) ? 2 : 3));
		rank = atomicAdd(&s_n[cls], 1);                         /* place inside the block's share of the class */
	}
	__syncthreads();
	if (tid <= PB_LEN_CLASSES && s_n[tid] > 0) {                /* one global atomic per class and block */
		s_base[tid] = atomicAdd(tid < PB_LEN_CLASSES ? &counts[tid] : general_count, s_n[tid]);
		if (tid == PB_LEN_CLASSES)
			atomicAdd(general_total, (unsigned long long) s_n[tid]);
	}
	__syncthreads();
	if (cls >= 0) {
		int *dst = cls < PB_LEN_CLASSES ? lists + (size_t) cls * cap : general_list;
		dst[s_base[cls] + rank] = pair;
	}
}

template <int ML, bool OVER, int WARPS_PER_BLOCK> constexpr size_t assemble_smem_bytes() {
	constexpr size_t LUT_BYTES = (2 * PB_NQM * PB_NQM + 2 + (OVER ? 2 * PB_NQ * PB_NQ : 0) + 2 * PB_NQM) * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned) + 512 * sizeof(uint16_t) + 2 * (PB_MAX_LEN + 2);
	constexpr size_t LUT_ALIGNED = (LUT_BYTES + 127) & ~(size_t) 127;
	return LUT_ALIGNED + sizeof(WarpSmem<ML>) * WARPS_PER_BLOCK;
}

/* ---- pack: flat AoS panda_qual -> packed records (one warp per pair) -------------- */
__global__ void pack_kernel(int n, const uint8_t *__restrict__ f_data, const unsigned long long *__restrict__ f_off, unsigned long long f_base,
                            const uint8_t *__restrict__ r_data, const unsigned long long *__restrict__ r_off, unsigned long long r_base,
                            const uint32_t *__restrict__ rec_off16, uint8_t *__restrict__ reads, pb_pair_meta *__restrict__ meta) {
	const int lane = threadIdx.x & 31;
	const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (pair >= n)
		return;
	/* offsets are absolute; f_data / r_data start at element f_base / r_base (a chunk of a larger batch) */
	const unsigned long long fb = f_off[pair] - f_base, rb = r_off[pair] - r_base;
	const int F = (int) (f_off[pair + 1] - f_base - fb), R = (int) (r_off[pair + 1] - r_base - rb);
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

/* records the caller packed into one buffer, shipped chunk by chunk: the chunk's offsets start at its first record */
__global__ void rebase_meta_kernel(int n, pb_pair_meta *__restrict__ meta, unsigned base16) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		meta[i].off16 -= base16;
}

}  // namespace pb
