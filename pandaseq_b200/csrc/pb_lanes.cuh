/* pb_lanes.cuh -- sm_100a kernel for the common case of the pair-assembly hot path: one LANE per read pair (kernel v4).
 *
 * assemble_kernel (pb_kernels.cuh) gives a whole warp to a pair; every step then pays for partly filled lane rounds,
 * shuffles, ballots and warp-wide bookkeeping.  Here a warp takes 32 consecutive pairs and every lane runs the
 * reference's align() (assembler.c:48-250) for its own pair as straight scalar code over 32-bit words of packed
 * bases.  What makes that affordable is that the rare cases do not have to be handled here at all: a lane that meets
 * one (a base that is not A/C/G/T, a quality outside 0..46, no seed at all, an overlap longer than a read, a read
 * outside this length class) appends its pair to a deferral list and the exact general kernel assembles those pairs
 * in a second launch on the same stream.  Both kernels implement the same function, so the split is invisible in the
 * results.
 *
 *   stage   32 bulk async copies (one per lane, cp.async.bulk / UBLKCP) land the 32 records in this warp's shared
 *           memory at a stride of an odd number of 16-byte units; one mbarrier per warp.
 *   seed    K1-K3 (assembler.c:92-118): the lane streams its forward read, forms the 2-bit digit of every base
 *           (misc.h:41) eight at a time, and inserts each 8-mer into its own 256-slot open-addressing table
 *           (lane-interleaved in shared memory, so 32 lanes always hit 32 different banks).  A slot is
 *           code:16 | first position:8 | second position:8 -- exactly what the reference's 65536x2 table remembers
 *           of a code ("first two positions", assembler.c:93-100).  The reverse read (template order) then probes;
 *           every hit sets a bit of the lane's candidate mask (BIT_LIST_SET, assembler.c:108).
 *   score   K4/K5 (assembler.c:120-143): candidates in increasing overlap; the count-based scorers
 *           (algo_simple_bayes.c:33-66, algo_uparse.c:33-66, algo_flash.c:30-60) need one AND + POPC per 8 bases.
 *   merge   K6 (assembler.c:158-244): merged bases 8 per word; the per-base posterior is summed in the reference's
 *           own order (forward-only stretch, overlap, reverse-only stretch, each left to right), so `quality` is
 *           bit-identical, not merely within tolerance.
 *
 * Configurations this kernel takes (the host decides, pb_device.cu): simple_bayesian / uparse / flash, no primers, no
 * trims, no overhang trimmer, no per-base log p requested, filters that read only the result record, reads <= 160 nt.
 */
#pragma once
#include "pb_kernels.cuh"

namespace pbl {

using pb::FULL;
using pb::NIB1;

constexpr int TRI = PB_NQM * (PB_NQM + 1) / 2;      /* entries of one triangular posterior table */
constexpr int TRI47 = PB_NQ * (PB_NQ + 1) / 2;      /* row 47 ("the other read is absent or masked"): qual_score[] */
constexpr uint8_t ST_DEFER = 255;

template <int ML> struct LaneArea {
	static constexpr int REC0 = (((ML + 7) / 8) * 4 * 2 + ((ML + 3) / 4) * 4 * 2 + 15) & ~15;
	static constexpr int REC_STRIDE = ((REC0 / 16) | 1) * 16;    /* odd number of 16-byte units: a lane's 128-bit loads never collide */
	static constexpr int SLOTS = 256;
	static constexpr int CW = (ML + 31) / 32;                    /* words of the candidate mask */
	static_assert(ML - 7 < 256, "positions are stored in 8 bits");
	static_assert(ML - 8 < SLOTS * 3 / 4, "the table must stay sparse");
	alignas(128) uint8_t rec[32 * REC_STRIDE];
	alignas(16) uint32_t tab[SLOTS * 32];                        /* slot s of lane l: tab[s * 32 + l] */
	uint32_t cmask[CW * 32];                                     /* word w of lane l: cmask[w * 32 + l] */
	alignas(8) uint64_t bar;
};

/* misc.h:41 on a word of eight one-hot bases: T=3 G=2 C=1 A=0, digit of base k in bits 4k, 4k+1 */
__device__ __forceinline__ unsigned digits8(unsigned x) {
	const unsigned x1 = x >> 1, x2 = x >> 2, x3 = x >> 3;
	const unsigned lo = (x1 | x3) & NIB1;
	const unsigned hi = (x2 | x3) & NIB1;
	return lo | (hi << 1);
}
/* eight nibble-spaced digits -> 16 bits (an injective packing; both reads use the same one) */
__device__ __forceinline__ unsigned code16(unsigned c) {
	return (c | (c >> 14)) & 0xFFFFu;
}
__device__ __forceinline__ unsigned slot_of(unsigned code) {
	return (code * 0x9E3779B1u) >> 24;
}
/* some nibble of x is zero */
__device__ __forceinline__ unsigned zero_nib(unsigned x) {
	return (x - NIB1) & ~x & 0x88888888u;
}
__device__ __forceinline__ unsigned tri_index(unsigned a, unsigned b) {
	const unsigned lo = min(a, b), hi = max(a, b);
	return ((hi * hi + hi) >> 1) + lo;
}

template <int ML, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
assemble_lanes_kernel(const pb_device_params *__restrict__ prm, int n,
                      const uint8_t *__restrict__ reads, const pb_pair_meta *__restrict__ meta,
                      pb_pair_result *__restrict__ results, uint8_t *__restrict__ seq_nt, long long seq_stride,
                      unsigned long long *__restrict__ counters, int *__restrict__ defer_list, int *__restrict__ defer_count,
                      unsigned long long *__restrict__ defer_total) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	using LA = LaneArea<ML>;
	double *s_tri = reinterpret_cast<double *>(smem_raw);                 /* [2][TRI]: posterior by (match, max q, min q) */
	unsigned *s_cnt = reinterpret_cast<unsigned *>(s_tri + 2 * TRI);
	constexpr size_t HEAD = ((2 * TRI * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned)) + 127) & ~(size_t) 127;
	LA *areas = reinterpret_cast<LA *>(smem_raw + HEAD);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < 2 * PB_NQM * PB_NQM; i += blockDim.x) {
		const int m = i / (PB_NQM * PB_NQM), a = (i / PB_NQM) % PB_NQM, b = i % PB_NQM;
		if (b <= a)
			s_tri[m * TRI + a * (a + 1) / 2 + b] = prm->recon[m][a][b];
	}
	for (int i = tid; i < PB_NCOUNTERS; i += blockDim.x)
		s_cnt[i] = 0;
	LA &wa = areas[warp];
	for (int k = lane; k < LA::SLOTS * 32; k += 32)
		wa.tab[k] = 0;
	for (int k = lane; k < LA::CW * 32; k += 32)
		wa.cmask[k] = 0;
	if (lane == 0) {
		pb::mbar_init(&wa.bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const double qual_nn = prm->qual_nn, pmatch = prm->sb_pmatch, pmismatch = prm->sb_pmismatch, threshold = prm->threshold;
	const int mo = prm->minoverlap, cfg_maxov = prm->maxoverlap, algo = prm->algo, nf = prm->nfilters;
	const long long nt_row = seq_stride / 2;
	const int out_cap = (int) seq_stride;
	const int nbatch = (n + 31) >> 5;
	uint32_t *const tab = wa.tab + lane;
	uint32_t *const cm = wa.cmask + lane;
	const uint8_t *const rb = wa.rec + lane * LA::REC_STRIDE;
	unsigned parity = 0;

	for (int batch = blockIdx.x * WARPS + warp; batch < nbatch; batch += gridDim.x * WARPS) {
		const int pair = batch * 32 + lane;
		unsigned off16 = 0;
		int F = 0xFFFF, R = 0;
		if (pair < n) {
			const uint2 mraw = *reinterpret_cast<const uint2 *>(&meta[pair]);
			off16 = mraw.x;
			F = (int) (mraw.y & 0xFFFFu);
			R = (int) (mraw.y >> 16);
		}
		const bool skip = F == 0xFFFF;              /* not a pair (FASTQ reader, fastq.c:176), or past the end of the batch */
		bool defer = !skip && (F > ML || R > ML || F < 16 || R < 16 || mo >= min(F, R));
		const bool act = !skip && !defer;
		const unsigned bytes = act ? pb::record_bytes((unsigned) F, (unsigned) R) : 0u;
		const unsigned total = __reduce_add_sync(FULL, bytes);
		if (lane == 0)
			pb::mbar_expect_tx(&wa.bar, total);
		__syncwarp();
		if (bytes)
			pb::bulk_g2s(wa.rec + lane * LA::REC_STRIDE, reads + (size_t) off16 * 16, bytes, &wa.bar);
		{       /* the next batch of this warp: pull its records into L2 while this one is processed */
			const long long np = (long long) pair + (long long) gridDim.x * WARPS * 32;
			if (np < n) {
				const uint2 mn = *reinterpret_cast<const uint2 *>(&meta[np]);
				const unsigned nF = mn.y & 0xFFFFu, nR = mn.y >> 16;
				if (nF != 0xFFFFu && nF <= (unsigned) ML && nR <= (unsigned) ML) {
					const unsigned nbytes = pb::record_bytes(nF, nR);
					asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(reads + (size_t) mn.x * 16), "r"(nbytes) : "memory");
				}
			}
		}
		pb::mbar_wait(&wa.bar, parity);
		parity ^= 1u;

		uint8_t status = PB_PAIR_OK;
		int slow = 0, bestov = -1, examined = 0, seq_len = 0, mism = 0;
		double best = 0.0, quality = 0.0;
		const int fwb = ((F + 7) >> 3) * 4, rwb = ((R + 7) >> 3) * 4;
		const uint32_t *const fnt = reinterpret_cast<const uint32_t *>(rb);
		const uint32_t *const rnt = reinterpret_cast<const uint32_t *>(rb + fwb);
		const uint32_t *const fq32 = reinterpret_cast<const uint32_t *>(rb + fwb + rwb);
		const uint32_t *const rq32 = fq32 + ((F + 3) >> 2);
		const uint8_t *const fq8 = reinterpret_cast<const uint8_t *>(fq32);
		const uint8_t *const rq8 = reinterpret_cast<const uint8_t *>(rq32);
		int maxov = 0, nbits = 0;

		if (act) {
			/* assembler.c:59,78-84 */
			maxov = cfg_maxov == 0 ? min(F, R) : min(F + R - mo - 1, cfg_maxov);
			nbits = (mo <= maxov) ? (maxov - mo + 1) : 1;
			if (nbits > LA::CW * 32)
				defer = true;
		}
		if (act && !defer) {
			/* ---- K1: forward 8-mers into this lane's table (assembler.c:92-103) ---- */
			unsigned pops = 0, zf = 0;
			{
				const int nw = (F + 7) >> 3;
				unsigned x = fnt[0];
				pops = __popc(x);
				zf = zero_nib(x);
				unsigned dprev = digits8(x);
				for (int w = 1; w < nw; w++) {
					x = fnt[w];
					pops += __popc(x);
					const int inw = min(F - 8 * w, 8);
					zf |= zero_nib(x | ~pb::nibmask(inw));
					const unsigned d = digits8(x);
#pragma unroll
					for (int t = 0; t < 8; t++) {
						if (t < inw) {
							const unsigned c = (t == 7) ? d : __funnelshift_r(dprev, d, 4 * (t + 1));
							const unsigned code = code16(c);
							const unsigned pp = (unsigned) (8 * w + t - 7);     /* position p = 8w+t, stored as p-7 (1..) */
							const unsigned key = code << 16;
							unsigned s = slot_of(code);
							for (;;) {
								const unsigned e = tab[s * 32];
								if (e == 0u) {
									tab[s * 32] = key | (pp << 8);
									break;
								}
								if ((e ^ key) < 0x10000u) {
									if ((e & 0xFFu) == 0u)
										tab[s * 32] = e | pp;                   /* the second position of this code; later ones are lost */
									break;
								}
								s = (s + 1) & (LA::SLOTS - 1);
							}
						}
					}
					dprev = d;
				}
				if (pops != (unsigned) F || zf != 0u)
					defer = true;                       /* some base is not exactly one of A, C, G, T */
			}
			/* ---- K2: reverse 8-mers probe (assembler.c:104-112); template order, so overlap = F - p + e ---- */
			{
				const int nw = (R + 7) >> 3;
				const int cb = F - mo - 7;              /* index = F - mo - p + e, p = stored + 7 */
				unsigned last = 0xFFFFFFFFu;
				unsigned x = rnt[0];
				pops = __popc(x);
				zf = zero_nib(x);
				unsigned dprev = digits8(x);
				for (int w = 1; w < nw; w++) {
					x = rnt[w];
					pops += __popc(x);
					const int inw = min(R - 8 * w, 8);
					zf |= zero_nib(x | ~pb::nibmask(inw));
					const unsigned d = digits8(x);
#pragma unroll
					for (int t = 0; t < 8; t++) {
						if (t < inw) {
							const unsigned c = (t == 7) ? d : __funnelshift_r(dprev, d, 4 * (t + 1));
							const unsigned code = code16(c);
							const unsigned key = code << 16;
							unsigned s = slot_of(code);
							unsigned e;
							for (;;) {
								e = tab[s * 32];
								if (e == 0u || (e ^ key) < 0x10000u)
									break;
								s = (s + 1) & (LA::SLOTS - 1);
							}
							if (e != 0u) {
								const unsigned base = (unsigned) (cb + 8 * w + t);
								const unsigned i1 = base - ((e >> 8) & 0xFFu);
								if (i1 != last && i1 < (unsigned) nbits) {
									last = i1;
									cm[(i1 >> 5) * 32] |= 1u << (i1 & 31u);
								}
								const unsigned p2 = e & 0xFFu;
								if (p2 != 0u) {
									const unsigned i2 = base - p2;
									if (i2 < (unsigned) nbits) {
										last = i2;
										cm[(i2 >> 5) * 32] |= 1u << (i2 & 31u);
									}
								}
							}
						}
					}
					dprev = d;
				}
				if (pops != (unsigned) R || zf != 0u)
					defer = true;
			}
		}
		/* ---- K3: clear (assembler.c:113-116), all lanes together ---- */
		__syncwarp();
		{
			const uint4 z = make_uint4(0, 0, 0, 0);
			uint4 *t4 = reinterpret_cast<uint4 *>(wa.tab);
#pragma unroll 8
			for (int k = 0; k < LA::SLOTS * 32 * 4 / 16 / 32; k++)
				t4[k * 32 + lane] = z;
		}
		__syncwarp();
		if (act) {
			/* ---- K4/K5: the candidates in increasing overlap (assembler.c:118-143) ---- */
			unsigned any = 0;
			best = qual_nn * (double) (unsigned long long) (F + R);          /* assembler.c:60 */
#pragma unroll 1
			for (int w = 0; w < LA::CW; w++) {
				unsigned mw = cm[w * 32];
				cm[w * 32] = 0;
				if (defer)
					mw = 0;
				any |= mw;
				while (mw) {
					const int ov = w * 32 + __ffs(mw) - 1 + mo;
					mw &= mw - 1;
					examined++;
					if (ov > F || ov > R) {              /* only with an explicit maxoverlap: left to the general kernel */
						defer = true;
						continue;
					}
					/* forward base F-ov+i against template-order reverse base i, i in [0, ov) */
					const int fs = F - ov, nw = (ov + 7) >> 3, sh = (fs & 7) * 4;
					const uint32_t *fp = fnt + (fs >> 3);
					unsigned lo = fp[0];
					int matches = 0;
					for (int k = 0; k < nw; k++) {
						const unsigned hi = fp[k + 1];
						const unsigned f = __funnelshift_r(lo, hi, sh);
						lo = hi;
						unsigned mt = f & rnt[k];
						if (k == nw - 1)
							mt &= pb::nibmask(ov - 8 * k);
						matches += __popc(mt);
					}
					const int mm = ov - matches;
					double prob;
					if (algo == PB_FLASH) {
						/* algo_flash.c:59: integer division inside log() */
						prob = (mm == ov) ? 0.0 : -CUDART_INF;
					} else {
						/* algo_simple_bayes.c:61-65 / algo_uparse.c:61-65 */
						const unsigned long long nn_count = (unsigned long long) ((long long) F + R - 2 * (long long) ov);
						prob = qual_nn * (double) nn_count + (double) matches * pmatch;
						prob = prob + (double) mm * pmismatch;
					}
					if (prob > best) {                   /* strict, ascending overlap: assembler.c:128-131 */
						best = prob;
						bestov = ov;
					}
				}
			}
			if (any == 0u)
				defer = true;                           /* no seed at all: every overlap is scored (assembler.c:118), the general kernel's job */
			if ((long long) examined == (long long) maxov - mo + 1)    /* assembler.c:135-137 */
				slow = 1;
			if (!defer && bestov < 0)
				status = PB_PAIR_NOALGN;
		}
		if (act && !defer && status == PB_PAIR_OK) {
			/* ---- K6: reconstruction (assembler.c:145-250) with forward_offset = reverse_offset = 0 ---- */
			const int df = F - bestov, dr = R - bestov;
			seq_len = df + R;
			const int len = seq_len + 1;
			/* B-cliff (assembler.c:176-177): trailing run of quality 2 in each read */
			int unmasked_f = F, lead_r = 0;
			if (fq8[F - 1] == 2) {
				while (unmasked_f > 0 && fq8[unmasked_f - 1] == 2)
					unmasked_f--;
			}
			if (rq8[0] == 2) {
				while (lead_r < R && rq8[lead_r] == 2)
					lead_r++;
			}
			unsigned qbad = 0;
			const double *const score = s_tri + TRI47;
			/* forward-only stretch: positions [0, df) (assembler.c:162-173) */
			double fquality = 0.0;
			{
				const int nwq = (df + 3) >> 2;
				for (int w = 0; w < nwq; w++) {
					const unsigned q4 = fq32[w];
					qbad |= ((q4 & 0x7F7F7F7Fu) + 0x51515151u) | q4;
					const int nb = min(df - 4 * w, 4);
#pragma unroll
					for (int t = 0; t < 4; t++)
						if (t < nb)
							fquality += score[(q4 >> (8 * t)) & 0x3Fu];
				}
			}
			/* overlap: position df+i pairs forward base df+i with template-order reverse base i (assembler.c:181-228) */
			double oquality = 0.0;
			{
				const int nwq = (bestov + 3) >> 2;
				const int shq = (df & 3) * 8, shn = (df & 7) * 4;
				const uint32_t *fqp = fq32 + (df >> 2);
				const uint32_t *fnp = fnt + (df >> 3);
				unsigned qlo = fqp[0];
				unsigned mbits = 0;
				for (int w = 0; w < nwq; w++) {
					const unsigned qhi = fqp[w + 1];
					const unsigned qa4 = __funnelshift_r(qlo, qhi, shq);
					qlo = qhi;
					const unsigned qb4 = rq32[w];
					qbad |= ((qb4 & 0x7F7F7F7Fu) + 0x51515151u) | qb4;
					if ((w & 1) == 0) {
						const int k = w >> 1;
						const unsigned f = __funnelshift_r(fnp[k], fnp[k + 1], shn);
						mbits = f & rnt[k];               /* one-hot bases: a nibble is non-zero iff the bases match */
					}
					const int nb = min(bestov - 4 * w, 4);
#pragma unroll
					for (int t = 0; t < 4; t++) {
						if (t < nb) {
							const int i = 4 * w + t;
							unsigned a = (qa4 >> (8 * t)) & 0x3Fu, b = (qb4 >> (8 * t)) & 0x3Fu;
							if (df + i >= unmasked_f)
								a = PB_NQ;
							if (i < lead_r)
								b = PB_NQ;
							const unsigned isnz = ((mbits >> (4 * (i & 7))) & 15u) != 0u ? (unsigned) TRI : 0u;
							oquality += s_tri[isnz + tri_index(a, b)];
						}
					}
				}
				/* forward qualities of the overlap were not range-checked above: positions [df, F) */
				for (int w = df >> 2; w < ((F + 3) >> 2); w++) {
					const unsigned q4 = fq32[w];
					qbad |= ((q4 & 0x7F7F7F7Fu) + 0x51515151u) | q4;
				}
			}
			/* reverse-only stretch: template-order reverse bases [bestov, R) (assembler.c:231-243) */
			double rquality = 0.0;
			{
				const int shq = (bestov & 3) * 8;
				const uint32_t *rqp = rq32 + (bestov >> 2);
				const int nwq = (dr + 3) >> 2;
				unsigned qlo = rqp[0];
				for (int w = 0; w < nwq; w++) {
					const unsigned qhi = rqp[w + 1];
					unsigned q4 = __funnelshift_r(qlo, qhi, shq);
					qlo = qhi;
					const int nb = min(dr - 4 * w, 4);
					if (nb < 4)
						q4 &= (1u << (8 * nb)) - 1u;      /* what follows the read in shared memory is not a quality */
					qbad |= ((q4 & 0x7F7F7F7Fu) + 0x51515151u) | q4;
#pragma unroll
					for (int t = 0; t < 4; t++)
						if (t < nb)
							rquality += score[(q4 >> (8 * t)) & 0x3Fu];
				}
			}
			if (qbad & 0x80808080u)
				defer = true;                           /* a quality outside 0..46: PHREDCLAMP (prob.h:23) is the general kernel's job */
			quality = (fquality + rquality + oquality) / (double) len;      /* assembler.c:244: divides by len, not seq_len */

			/* merged bases, eight per word */
			uint8_t *const orow = seq_nt ? seq_nt + (size_t) pair * nt_row : nullptr;
			const int nwords = (seq_len + 7) >> 3;
			unsigned prev = 0;
			for (int k = 0; k < nwords; k++) {
				const int idx0 = 8 * k;
				const int nV = min(seq_len - idx0, 8);
				const int nF = min(max(F - idx0, 0), 8);
				const int r0 = min(max(df - idx0, 0), 8);
				const unsigned maskV = pb::nibmask(nV), maskF = pb::nibmask(nF) & maskV, maskR = maskV & ~pb::nibmask(r0);
				const unsigned fw = (idx0 < F ? fnt[k] : 0u) & maskF;
				const unsigned rw = pb::nibwin(rnt, idx0 - df) & maskR;
				const unsigned both = maskF & maskR;
				const unsigned andw = fw & rw;
				unsigned missb = both & NIB1 & ~pb::nz_nib(andw);
				unsigned nt = (fw & ~maskR) | (rw & ~maskF) | andw | (fw & (missb * 15u));
				mism += __popc(missb);
				while (missb) {                          /* assembler.c:215-219: the strictly better quality wins, ties go to the forward read */
					const int t = (__ffs(missb) - 1) >> 2;
					missb &= missb - 1;
					if (fq8[idx0 + t] < rq8[idx0 - df + t])
						nt = (nt & ~(15u << (4 * t))) | (rw & (15u << (4 * t)));
				}
				if (k & 1) {
					if (orow && idx0 - 8 < out_cap)
						*reinterpret_cast<uint2 *>(orow + (size_t) (k - 1) * 4) = make_uint2(prev, nt);
				} else {
					prev = nt;
				}
			}
			if ((nwords & 1) && orow && 8 * (nwords - 1) < out_cap)
				*reinterpret_cast<uint2 *>(orow + (size_t) (nwords - 1) * 4) = make_uint2(prev, 0u);

			if (!defer) {
				if (quality < threshold) {               /* assembler.c:334-338 */
					status = PB_PAIR_LOWQ;
				} else {
					/* module_checkseq (module.c:124-137): the first failing check rejects the pair */
					for (int k = 0; k < nf; k++) {
						const int kind = prm->filters[k].kind, iv = prm->filters[k].ivalue;
						bool pass = true;
						if (kind == PB_FILTER_SHORT)
							pass = seq_len >= iv;
						else if (kind == PB_FILTER_LONG)
							pass = seq_len <= iv;
						else if (kind == PB_FILTER_MIN_OVERLAPBITS)
							pass = prm->filters[k].dvalue * 0.693147180559945309417232121458 <= best;
						else if (kind == PB_FILTER_MISS_THE_POINT)
							pass = mism <= iv;
						/* PB_FILTER_NO_N: reads of A/C/G/T only merge into A/C/G/T only */
						if (!pass) {
							status = (uint8_t) (PB_PAIR_FILTERED + k);
							break;
						}
					}
				}
			}
		}
		/* ---- result records and counters ---- */
		if (defer)
			status = ST_DEFER;
		else if (skip)
			status = PB_PAIR_SKIP;
		if (pair < n && status != ST_DEFER) {
			union { pb_pair_result r; uint4 v[2]; } ru;
			ru.v[0] = make_uint4(0, 0, 0, 0);
			ru.v[1] = make_uint4(0, 0, 0, 0);
			ru.r.status = status;
			if (status != PB_PAIR_SKIP) {
				ru.r.slow = (uint8_t) slow;
				ru.r.examined = (uint16_t) examined;
				if (status != PB_PAIR_NOALGN) {
					ru.r.overlap = (uint16_t) bestov;
					ru.r.seq_len = (uint16_t) seq_len;
					ru.r.mismatches = (uint16_t) mism;
					ru.r.quality = quality;
					ru.r.est_prob = best;
				}
			}
			uint4 *dst = reinterpret_cast<uint4 *>(&results[pair]);
			dst[0] = ru.v[0];
			dst[1] = ru.v[1];
		}
		__syncwarp();
		const bool counted = status != ST_DEFER && status != PB_PAIR_SKIP;
		const unsigned m_count = __ballot_sync(FULL, counted);
		const unsigned m_ok = __ballot_sync(FULL, status == PB_PAIR_OK);
		const unsigned m_lowq = __ballot_sync(FULL, status == PB_PAIR_LOWQ);
		const unsigned m_noalgn = __ballot_sync(FULL, status == PB_PAIR_NOALGN);
		const unsigned m_slow = __ballot_sync(FULL, counted && slow);
		const unsigned m_defer = __ballot_sync(FULL, status == ST_DEFER);
		const unsigned longest = __reduce_max_sync(FULL, status == PB_PAIR_OK ? (unsigned) bestov : 0u);
		if (status == PB_PAIR_OK)
			atomicAdd(&s_cnt[PB_C_OVERLAPS + bestov], 1u);
		else if (counted && status >= PB_PAIR_FILTERED)
			atomicAdd(&s_cnt[PB_C_REJECTED + status - PB_PAIR_FILTERED], 1u);
		int dbase = 0;
		if (lane == 0) {
			if (m_count) atomicAdd(&s_cnt[PB_C_COUNT], (unsigned) __popc(m_count));
			if (m_ok) atomicAdd(&s_cnt[PB_C_OK], (unsigned) __popc(m_ok));
			if (m_lowq) atomicAdd(&s_cnt[PB_C_LOWQ], (unsigned) __popc(m_lowq));
			if (m_noalgn) atomicAdd(&s_cnt[PB_C_NOALGN], (unsigned) __popc(m_noalgn));
			if (m_slow) atomicAdd(&s_cnt[PB_C_SLOW], (unsigned) __popc(m_slow));
			if (m_ok) atomicMax(&s_cnt[PB_C_LONGEST], longest);
			if (m_defer) {
				dbase = atomicAdd(defer_count, __popc(m_defer));
				atomicAdd(defer_total, (unsigned long long) __popc(m_defer));       /* diagnostics: pb_lanes_stats() */
			}
		}
		if (m_defer) {
			dbase = __shfl_sync(FULL, dbase, 0);
			if (status == ST_DEFER)
				defer_list[dbase + __popc(m_defer & pb::lanemask_lt())] = pair;
		}
		__syncwarp();      /* every lane is done with its record before the next batch lands on it */
	}
	__syncthreads();
	for (int i = tid; i < PB_NCOUNTERS; i += blockDim.x) {
		const unsigned c = s_cnt[i];
		if (c) {
			if (i == PB_C_LONGEST)
				atomicMax(&counters[i], (unsigned long long) c);
			else
				atomicAdd(&counters[i], (unsigned long long) c);
		}
	}
}

template <int ML, int WARPS> constexpr size_t lanes_smem_bytes() {
	constexpr size_t HEAD = ((2 * TRI * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned)) + 127) & ~(size_t) 127;
	return HEAD + sizeof(LaneArea<ML>) * WARPS;
}

}  // namespace pbl
