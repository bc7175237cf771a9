/* pb_lanes.cuh -- second half of the two-kernel path for the common configurations: one LANE per read pair.
 *
 * assemble_kernel (pb_kernels.cuh) gives a whole warp to a pair from the record to the result; after the k-mer join
 * (K1-K3, which does want 32 lanes on one pair) everything that is left -- scoring a handful of candidate overlaps,
 * merging the reads, summing the per-base posterior -- is a stream of 32-bit words of packed bases, and a warp-wide
 * formulation pays for partly filled lane rounds, shuffles, ballots and per-pair bookkeeping.  So the common
 * configurations run as two kernels: the seeding kernel (pbs::sweep_seed_kernel, lane per pair, or pb::seed_kernel, warp per
 * pair) leaves the candidate overlaps of every pair as a bit mask, and the kernel here takes 32 pairs of one overlap bin per warp, every lane running K4-K6 of the reference's
 * align() (assembler.c:118-250) for its own pair as straight scalar code.  What makes that affordable is that the rare
 * cases do not have to be handled here: a lane that meets one (a base that is not A/C/G/T, a quality outside 0..46, no
 * seed at all, an overlap longer than a read, a read shorter than 16 nt) appends its pair to a deferral list and the
 * exact general kernel assembles those pairs in a third launch on the same stream.  All kernels implement the same
 * function, so the split is invisible in the results.
 *
 *   stage   32 bulk async copies (one per lane, cp.async.bulk / UBLKCP) land the 32 records in this warp's shared
 *           memory at a stride of an odd number of 16-byte units; one mbarrier per warp; the warp's next batch (its
 *           number, pairs, metadata and seeds records) is fetched one batch ahead and its records are prefetched into L2
 *           (cp.async.bulk.prefetch / UBLKPF) meanwhile.
 *   score   K4/K5 (assembler.c:120-143): candidates in increasing overlap; the count-based scorers
 *           (algo_simple_bayes.c:33-66, algo_uparse.c:33-66, algo_flash.c:30-60) need one AND + POPC per 8 bases, pear
 *           (algo_pear.c:32-59) one table gather per base, added in the reference's order.
 *   merge   K6 (assembler.c:158-244): merged bases 8 per word (copy of the forward words / both reads / shifted copy of the
 *           reverse words); the per-base posterior is summed per stretch (forward-only, overlap, reverse-only) as the
 *           reference does, on two accumulators each, the two outer stretches in one loop, the overlap eight bases per step.
 *
 * A first version of this file also did the k-mer join per lane (a 256-slot open-addressing table per lane, lane-
 * interleaved in shared memory).  Measured on B200: 124 Mpairs/s against 650 for the warp-per-pair kernel -- 1 KB of
 * table per pair leaves 4 warps per SM, and the probe chains of 32 lanes are as long as the longest of them
 * (1,541 warp-instructions per pair at 13 active lanes, 0.17 instructions per cycle and scheduler).  DESIGN.md section 5.
 *
 * Configurations this path takes (the host decides, pb_device.cu): simple_bayesian / uparse / flash / pear, no primers, no
 * trims, no overhang trimmer, no per-base log p requested, filters that read only the result record; length classes 152, 160,
 * 256 and 320 nt (12, 11, 7 and 6 warps per SM); a batch of mixed lengths is listed by class first (pb::class_list_kernel) and
 * every class runs the kernel sized for it, pairs with a read above 320 nt go to the general kernel.
 */
#pragma once
#include "pb_kernels.cuh"

namespace pbl {

using pb::FULL;
using pb::NIB1;
using pb::pear_test_pass;

constexpr int LUT_DOUBLES = 2 * PB_NQM * PB_NQM + 256;      /* the posterior table, then qual_score[] padded to one entry per byte value */
constexpr uint8_t ST_DEFER = 255;

template <int ML> struct LaneArea {
	static constexpr int REC0 = (((ML + 7) / 8) * 4 * 2 + ((ML + 3) / 4) * 4 * 2 + 15) & ~15;
	static constexpr int REC_STRIDE = ((REC0 / 16) | 1) * 16;    /* odd number of 16-byte units: a lane's 128-bit loads never collide */
	static constexpr int CW = pb::seed_mask_words(ML);           /* words of the candidate mask (pb::seed_kernel) */
	static constexpr int SWORDS = pb::seed_words(ML);
	static_assert(ML <= 320, "the bins of the seeds record reach up to 320 nt");
	alignas(128) uint8_t rec[32 * REC_STRIDE];
	alignas(8) uint64_t bar;
};

template <int ML, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
assemble_lanes_kernel(const pb_device_params *__restrict__ prm, int n,
                      const uint8_t *__restrict__ reads, const pb_pair_meta *__restrict__ meta,
                      const uint32_t *__restrict__ seeds, const int *__restrict__ order, pb_pair_result *__restrict__ results, uint8_t *__restrict__ seq_nt, long long seq_stride,
                      unsigned long long *__restrict__ counters, int *__restrict__ defer_list, int *__restrict__ defer_count,
                      unsigned long long *__restrict__ defer_total, unsigned *__restrict__ next_batch, const int *__restrict__ order_n, int pair_base) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	using LA = LaneArea<ML>;
	double *s_rec = reinterpret_cast<double *>(smem_raw);                 /* recon[2][48][48], row = quality a + 48 * match */
	double *s_score = s_rec + LUT_DOUBLES - 256;                          /* qual_score[] by raw byte: 256 entries, [47..] zero */
	unsigned *s_cnt = reinterpret_cast<unsigned *>(s_rec + LUT_DOUBLES);
	constexpr size_t HEAD = ((LUT_DOUBLES * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned)) + 127) & ~(size_t) 127;
	LA *areas = reinterpret_cast<LA *>(smem_raw + HEAD);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < 2 * PB_NQM * PB_NQM; i += blockDim.x)
		s_rec[i] = (&prm->recon[0][0][0])[i];
	for (int i = tid; i < 256; i += blockDim.x)
		s_score[i] = i < PB_NQ ? prm->score[i] : 0.0;
	for (int i = tid; i < PB_NCOUNTERS; i += blockDim.x)
		s_cnt[i] = 0;
	LA &wa = areas[warp];
	if (lane == 0) {
		pb::mbar_init(&wa.bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const double qual_nn = prm->qual_nn, pmatch = prm->sb_pmatch, pmismatch = prm->sb_pmismatch, threshold = prm->threshold;
	const int mo = prm->minoverlap, cfg_maxov = prm->maxoverlap, algo = prm->algo, nf = prm->nfilters;
	const long long nt_row = seq_stride / 2;
	const int out_cap = (int) seq_stride;
	if (order_n)
		n = *order_n;         /* `order` holds the pairs of one length class (pb::class_list_kernel) */
	const int nbatch = (n + 31) >> 5;
	const uint8_t *const rb = wa.rec + lane * LA::REC_STRIDE;
	unsigned parity = 0;

	/* Which 32 entries of the bin list a warp takes next comes from a counter in global memory (*next_batch, zero at launch): with a
	 * fixed share per warp the warps the scheduler favours finish early and leave the SM half empty at the end.  Everything a batch
	 * needs before its records can be fetched -- its number (an atomic), its pairs (the bin list), their metadata and seeds records --
	 * is a chain of dependent global loads; it runs ahead of the batch in flight, every link one whole iteration after the one it
	 * depends on, so that no iteration waits for any of it: at the top of an iteration the counter is bumped (read at the top of the
	 * next one: three batches ahead), the pairs of the batch three ahead are looked up in the bin list, the metadata and seeds records
	 * of the batch two ahead are loaded, and the records of the next batch are pulled into L2. */
	struct Head {
		int pair;
		unsigned off16;
		int F, R;
		unsigned cw[LA::CW], sflags;
	};
	auto take = [&]() {
		int b = 0;
		if (lane == 0)
			b = (int) atomicAdd(next_batch, 1u);
		return b;                                              /* lane 0's value counts: __shfl_sync(FULL, b, 0) where it is needed */
	};
	auto pair_of = [&](int b) {
		const long long item = (long long) b * 32 + lane;
		return (b < nbatch && item < n) ? order[item] : -1;    /* pairs come bin by bin (pb::bin_order_kernel) */
	};
	auto load_head = [&](int pair, Head &h) {
		h.pair = pair;
		h.off16 = 0;
		h.F = 0xFFFF;                                          /* not a pair (FASTQ reader, fastq.c:176), or past the end of the list */
		h.R = 0;
		h.sflags = pb::PB_SEED_SKIP;
#pragma unroll
		for (int w = 0; w < LA::CW; w++)
			h.cw[w] = 0;
		if (pair >= 0) {
			const uint2 mraw = *reinterpret_cast<const uint2 *>(&meta[pair]);
			unsigned sw[LA::SWORDS];
#pragma unroll
			for (int q = 0; q < LA::SWORDS / 4; q++) {
				const uint4 v = reinterpret_cast<const uint4 *>(seeds)[(size_t) pair * (LA::SWORDS / 4) + q];
				sw[4 * q] = v.x; sw[4 * q + 1] = v.y; sw[4 * q + 2] = v.z; sw[4 * q + 3] = v.w;
			}
			h.off16 = mraw.x;
			h.F = (int) (mraw.y & 0xFFFFu);
			h.R = (int) (mraw.y >> 16);
#pragma unroll
			for (int w = 0; w < LA::CW; w++)
				h.cw[w] = sw[w];
			h.sflags = sw[LA::CW];
		}
	};
	int batch = __shfl_sync(FULL, take(), 0);
	int batch1 = __shfl_sync(FULL, take(), 0);
	int batch2 = __shfl_sync(FULL, take(), 0);
	int taken = take();                                        /* the third batch from here; read one iteration later */
	Head cur, nxt, far;
	load_head(pair_of(batch), cur);
	load_head(pair_of(batch1), nxt);
	int pair2 = pair_of(batch2);
	while (batch < nbatch) {
		const int batch3 = __shfl_sync(FULL, taken, 0);        /* asked for one iteration ago */
		taken = take();
		const int pair3 = pair_of(batch3);                     /* used one iteration from here */
		const int pair = cur.pair;
		const int F = cur.F, R = cur.R;
		unsigned cw[LA::CW];
#pragma unroll
		for (int w = 0; w < LA::CW; w++)
			cw[w] = cur.cw[w];
		const unsigned sflags = cur.sflags;
		const bool skip = F == 0xFFFF;
		bool defer = !skip && ((sflags & pb::PB_SEED_GENERAL) != 0 || F > ML || R > ML);      /* a record must fit its slot */
		const bool act = !skip && !defer;
		const unsigned bytes = act ? pb::record_bytes((unsigned) F, (unsigned) R) : 0u;
		const unsigned total = __reduce_add_sync(FULL, bytes);
		if (lane == 0)
			pb::mbar_expect_tx(&wa.bar, total);
		__syncwarp();
		if (bytes)
			pb::bulk_g2s(wa.rec + lane * LA::REC_STRIDE, reads + (size_t) cur.off16 * 16, bytes, &wa.bar);
		load_head(pair2, far);                                 /* the batch after next: its pairs were looked up one iteration ago */
		if (nxt.F != 0xFFFF && nxt.F <= ML && nxt.R <= ML) {      /* the next batch, known for an iteration: pull its records into L2 now */
			const unsigned nbytes = pb::record_bytes((unsigned) nxt.F, (unsigned) nxt.R);
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(reads + (size_t) nxt.off16 * 16), "r"(nbytes) : "memory");
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
		if (act) {
			/* ---- K4/K5: the candidates in increasing overlap (assembler.c:118-143) ---- */
			best = qual_nn * (double) (unsigned long long) (F + R);          /* assembler.c:60 */
#pragma unroll
			for (int w = 0; w < LA::CW; w++) {
				unsigned mw = defer ? 0u : cw[w];
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
					int matches = 0;
					if (algo != PB_PEAR) {               /* the count-based scorers: matching bases of the overlap */
						unsigned lo = fp[0];
						for (int k = 0; k < nw; k++) {
							const unsigned hi = fp[k + 1];
							const unsigned f = __funnelshift_r(lo, hi, sh);
							lo = hi;
							unsigned mt = f & rnt[k];
							if (k == nw - 1)
								mt &= pb::nibmask(ov - 8 * k);
							matches += __popc(mt);
						}
					}
					const int mm = ov - matches;
					double prob;
					if (algo == PB_PEAR) {
						/* algo_pear.c:32-59, term by term in its order.  Both qualities come from the FORWARD read (lines 52, 54
						 * index it with the reverse index); past its end that is the general kernel's defined case. */
						if (R > F) {
							defer = true;
							continue;
						}
						prob = 0.0;
						{
							/* four bases per step: forward qualities fs+i ascending, "reverse" qualities R-1-i descending (bytes of a
							 * window reversed), match bits from the packed bases, the match bit folded into the table row */
							const int nfull = ov >> 2;
							const int sha = (fs & 3) * 8, shb = ((R - 4) & 3) * 8, shn = (fs & 7) * 4;
							const uint32_t *qap = fq32 + (fs >> 2);
							const uint32_t *fnp = fnt + (fs >> 3);
							int cw0 = (R - 4) >> 2;                     /* word holding byte R-4-i0 */
							unsigned alo = qap[0], bhi = fq32[cw0 + 1], nlo = fnp[0], mw = 0;
							for (int w = 0; w < nfull; w++) {
								const unsigned ahi = qap[w + 1];
								const unsigned qa4 = __funnelshift_r(alo, ahi, sha) & 0x3F3F3F3Fu;
								alo = ahi;
								const unsigned blo = fq32[cw0 - w];
								const unsigned qb4 = __byte_perm(__funnelshift_r(blo, bhi, shb), 0, 0x0123) & 0x3F3F3F3Fu;
								bhi = blo;
								if ((w & 1) == 0) {
									const unsigned nhi = fnp[(w >> 1) + 1];
									mw = pb::nz_nib(__funnelshift_r(nlo, nhi, shn) & rnt[w >> 1]);
									nlo = nhi;
								} else {
									mw >>= 16;
								}
								unsigned sp = mw & 0x1111u;
								sp = (sp | (sp << 8)) & 0x00110011u;
								sp = (sp | (sp << 4)) & 0x01010101u;
								const unsigned row4 = qa4 + sp * 48u;
#pragma unroll
								for (int t = 0; t < 4; t++)
									prob += s_rec[__byte_perm(row4, 0, 0x4440 + t) * PB_NQM + __byte_perm(qb4, 0, 0x4440 + t)];
							}
							for (int i = 4 * nfull; i < ov; i++) {
								const int fi = fs + i;
								const unsigned fb = (fnt[fi >> 3] >> (4 * (fi & 7))) & 15u, rbase = (rnt[i >> 3] >> (4 * (i & 7))) & 15u;
								const unsigned qa = fq8[fi] & 0x3Fu, qb = fq8[R - 1 - i] & 0x3Fu;
								prob += s_rec[(((fb & rbase) ? PB_NQM : 0u) + qa) * PB_NQM + qb];
							}
						}
					} else if (algo == PB_FLASH) {
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
			if ((long long) examined == (long long) maxov - mo + 1)    /* assembler.c:135-137 */
				slow = 1;
			if (!defer && bestov < 0) {
				status = PB_PAIR_NOALGN;
				if (algo == PB_PEAR) {
					/* pear's scores above came from the low six bits of the raw qualities; a read with a quality outside 0..46
					 * (PHREDCLAMP, prob.h:23) may have lost its overlap to that, and only the merge below checks the range.
					 * Rare (no alignment): look at every quality of the pair before the verdict stands. */
					unsigned qb = 0;
					const int nq = ((F + 3) >> 2) + ((R + 3) >> 2);
					for (int w = 0; w < nq; w++) {
						const unsigned q4 = fq32[w];
						qb |= ((q4 & 0x7F7F7F7Fu) + 0x51515151u) | q4;
					}
					if (qb & 0x80808080u)
						defer = true;
				}
			}
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
			/* Range check of every quality that is added up (PHREDCLAMP, prob.h:23, is the general kernel's job): bit 7 of a byte of
			 * q + 0x51.. is set iff that quality is 47 or more, as long as no byte has its own bit 7 set (no carries between the
			 * bytes then) -- which the OR of the words themselves tells at the end. */
			unsigned qsum = 0, qor = 0;
			auto qcheck2 = [&](unsigned x, unsigned y) {
				qsum |= (x + 0x51515151u) | (y + 0x51515151u);
				qor |= x | y;
			};
			/* The sums below run on two accumulators each (even / odd bases): the additions are the reference's, their
			 * order is not, which moves `quality` by an ulp or two (tolerance 1e-6, BASELINE.json north_star).
			 * forward-only stretch: positions [0, df) (assembler.c:162-173); reverse-only stretch: template-order reverse bases
			 * [bestov, R) (assembler.c:231-243).  Two independent streams of qual_score[] gathers: they share one loop as far as both
			 * reach, which halves the loop overhead and lets the two chains of additions overlap. */
			double fquality = 0.0, fquality1 = 0.0, rquality = 0.0, rquality1 = 0.0;
			{
				const int shr = (bestov & 3) * 8;
				const uint32_t *rqp = rq32 + (bestov >> 2);
				const int nff = df >> 2, nfr = dr >> 2, nfc = min(nff, nfr);
				unsigned rlo = rqp[0];
				int w = 0;
				for (; w < nfc; w++) {
					const unsigned qf = fq32[w];
					const unsigned rhi = rqp[w + 1];
					const unsigned qr = __funnelshift_r(rlo, rhi, shr);
					rlo = rhi;
					qcheck2(qf, qr);
					fquality += s_score[__byte_perm(qf, 0, 0x4440)];
					rquality += s_score[__byte_perm(qr, 0, 0x4440)];
					fquality1 += s_score[__byte_perm(qf, 0, 0x4441)];
					rquality1 += s_score[__byte_perm(qr, 0, 0x4441)];
					fquality += s_score[__byte_perm(qf, 0, 0x4442)];
					rquality += s_score[__byte_perm(qr, 0, 0x4442)];
					fquality1 += s_score[__byte_perm(qf, 0, 0x4443)];
					rquality1 += s_score[__byte_perm(qr, 0, 0x4443)];
				}
				for (int v = w; v < nff; v++) {
					const unsigned qf = fq32[v];
					qcheck2(qf, 0u);
					fquality += s_score[__byte_perm(qf, 0, 0x4440)];
					fquality1 += s_score[__byte_perm(qf, 0, 0x4441)];
					fquality += s_score[__byte_perm(qf, 0, 0x4442)];
					fquality1 += s_score[__byte_perm(qf, 0, 0x4443)];
				}
				for (; w < nfr; w++) {
					const unsigned rhi = rqp[w + 1];
					const unsigned qr = __funnelshift_r(rlo, rhi, shr);
					rlo = rhi;
					qcheck2(qr, 0u);
					rquality += s_score[__byte_perm(qr, 0, 0x4440)];
					rquality1 += s_score[__byte_perm(qr, 0, 0x4441)];
					rquality += s_score[__byte_perm(qr, 0, 0x4442)];
					rquality1 += s_score[__byte_perm(qr, 0, 0x4443)];
				}
				/* the last one to three bases of each stretch; what lies beyond them in the word is checked where it belongs
				 * (the overlap) or is not a quality at all (what follows the read in shared memory) */
				const int nbf = df & 3, nbr = dr & 3;
				const unsigned qf = nbf ? fq32[nff] & ((1u << (8 * nbf)) - 1u) : 0u;
				const unsigned qr = nbr ? __funnelshift_r(rlo, rqp[nfr + 1], shr) & ((1u << (8 * nbr)) - 1u) : 0u;
				qcheck2(qf, qr);
				if (nbf > 0) fquality += s_score[__byte_perm(qf, 0, 0x4440)];
				if (nbr > 0) rquality += s_score[__byte_perm(qr, 0, 0x4440)];
				if (nbf > 1) fquality1 += s_score[__byte_perm(qf, 0, 0x4441)];
				if (nbr > 1) rquality1 += s_score[__byte_perm(qr, 0, 0x4441)];
				if (nbf > 2) fquality += s_score[__byte_perm(qf, 0, 0x4442)];
				if (nbr > 2) rquality += s_score[__byte_perm(qr, 0, 0x4442)];
				fquality += fquality1;
				rquality += rquality1;
			}
			/* overlap: position df+i pairs forward base df+i with template-order reverse base i (assembler.c:181-228), eight bases
			 * per step.  The posterior is recon[match][a][b]; the match bit is folded into the row index (row = a + 48 * match),
			 * B-cliff masked bases get index 47 (assembler.c:194-210).  The last step of a pair is the same code with the bases past
			 * the overlap zeroed and their additions predicated off, so that the lanes of a warp stay together. */
			double oquality = 0.0, oquality1 = 0.0;
			{
				const bool cliff = unmasked_f < F || lead_r > 0;
				const int ia = unmasked_f - df;           /* overlap bases i >= ia have a masked forward base */
				const int shq = (df & 3) * 8, shn = (df & 7) * 4;
				const uint32_t *fqp = fq32 + (df >> 2);
				const uint32_t *fnp = fnt + (df >> 3);
				unsigned qlo = fqp[0], nlo = fnp[0];
				const int nw8 = (bestov + 7) >> 3;
				for (int k = 0; k < nw8; k++) {
					const unsigned q1 = fqp[2 * k + 1], q2 = fqp[2 * k + 2];
					unsigned qa0 = __funnelshift_r(qlo, q1, shq), qa1 = __funnelshift_r(q1, q2, shq);
					qlo = q2;
					unsigned qb0 = rq32[2 * k], qb1 = rq32[2 * k + 1];
					const unsigned nhi = fnp[k + 1];
					const unsigned mw = pb::nz_nib(__funnelshift_r(nlo, nhi, shn) & rnt[k]);      /* bit 4t: bases t of this word match */
					nlo = nhi;
					const int nb = bestov - 8 * k;       /* bases of this step: 8, fewer in the last one */
					if (nb < 8) {                        /* what lies past the overlap must not reach the range check */
						const unsigned k0 = nb >= 4 ? 0xFFFFFFFFu : (1u << (8 * nb)) - 1u;
						const unsigned k1 = nb > 4 ? (1u << (8 * (nb - 4))) - 1u : 0u;
						qa0 &= k0; qb0 &= k0;
						qa1 &= k1; qb1 &= k1;
					}
					qcheck2(qa0, qb0);
					qcheck2(qa1, qb1);
					qa0 &= 0x3F3F3F3Fu;                  /* keeps the table index inside shared memory for pairs that are handed on */
					qb0 &= 0x3F3F3F3Fu;
					qa1 &= 0x3F3F3F3Fu;
					qb1 &= 0x3F3F3F3Fu;
					if (cliff) {
						const unsigned ma0 = __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - 8 * min(max(ia - 8 * k, 0), 4));
						const unsigned ma1 = __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - 8 * min(max(ia - 8 * k - 4, 0), 4));
						const unsigned mb0 = __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - 8 * min(max(lead_r - 8 * k, 0), 4));
						const unsigned mb1 = __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - 8 * min(max(lead_r - 8 * k - 4, 0), 4));
						qa0 = (qa0 & ma0) | (0x2F2F2F2Fu & ~ma0);
						qa1 = (qa1 & ma1) | (0x2F2F2F2Fu & ~ma1);
						qb0 = (qb0 & ~mb0) | (0x2F2F2F2Fu & mb0);
						qb1 = (qb1 & ~mb1) | (0x2F2F2F2Fu & mb1);
					}
					/* match bits 0,4,8,12 (16,20,24,28) -> bits 0,8,16,24, times 48, added to the forward qualities */
					unsigned s0 = mw & 0x1111u, s1 = (mw >> 16) & 0x1111u;
					s0 = (s0 | (s0 << 8)) & 0x00110011u;
					s1 = (s1 | (s1 << 8)) & 0x00110011u;
					s0 = (s0 | (s0 << 4)) & 0x01010101u;
					s1 = (s1 | (s1 << 4)) & 0x01010101u;
					const unsigned r0 = qa0 + s0 * 48u, r1 = qa1 + s1 * 48u;
					oquality += s_rec[__byte_perm(r0, 0, 0x4440) * PB_NQM + __byte_perm(qb0, 0, 0x4440)];
					if (nb > 1) oquality1 += s_rec[__byte_perm(r0, 0, 0x4441) * PB_NQM + __byte_perm(qb0, 0, 0x4441)];
					if (nb > 2) oquality += s_rec[__byte_perm(r0, 0, 0x4442) * PB_NQM + __byte_perm(qb0, 0, 0x4442)];
					if (nb > 3) oquality1 += s_rec[__byte_perm(r0, 0, 0x4443) * PB_NQM + __byte_perm(qb0, 0, 0x4443)];
					if (nb > 4) oquality += s_rec[__byte_perm(r1, 0, 0x4440) * PB_NQM + __byte_perm(qb1, 0, 0x4440)];
					if (nb > 5) oquality1 += s_rec[__byte_perm(r1, 0, 0x4441) * PB_NQM + __byte_perm(qb1, 0, 0x4441)];
					if (nb > 6) oquality += s_rec[__byte_perm(r1, 0, 0x4442) * PB_NQM + __byte_perm(qb1, 0, 0x4442)];
					if (nb > 7) oquality1 += s_rec[__byte_perm(r1, 0, 0x4443) * PB_NQM + __byte_perm(qb1, 0, 0x4443)];
				}
			}
			oquality += oquality1;
			if ((qsum | qor) & 0x80808080u)
				defer = true;                           /* a quality outside 0..46: PHREDCLAMP (prob.h:23) is the general kernel's job */
			quality = (fquality + rquality + oquality) / (double) len;      /* assembler.c:244: divides by len, not seq_len */

			/* merged bases, eight per word: the words that lie in the forward-only stretch are copies of the forward read's, the
			 * words past the forward read's end are a shifted window of the (template-order) reverse read; only the words in
			 * between see both reads */
			uint8_t *const orow = seq_nt ? seq_nt + (size_t) pair * nt_row : nullptr;
			const int nwords = (seq_len + 7) >> 3;
			unsigned prev = 0;
			auto emit = [&](int k, unsigned nt) {
				if (k & 1) {
					if (orow && 8 * (k - 1) < out_cap)
						*reinterpret_cast<uint2 *>(orow + (size_t) (k - 1) * 4) = make_uint2(prev, nt);
				} else {
					prev = nt;
				}
			};
			const int k1 = df >> 3, k2 = min((F + 7) >> 3, nwords);
			int k = 0;
			for (; k < k1; k++)
				emit(k, fnt[k]);
			for (; k < k2; k++) {
				const int idx0 = 8 * k;
				const int nV = min(seq_len - idx0, 8);
				const int nF = min(F - idx0, 8);
				const int r0 = min(max(df - idx0, 0), 8);
				const unsigned maskV = pb::nibmask(nV), maskF = pb::nibmask(nF) & maskV, maskR = maskV & ~pb::nibmask(r0);
				const unsigned fw = fnt[k] & maskF;
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
				emit(k, nt);
			}
			if (k < nwords) {
				const int pos = 8 * k - df;              /* >= overlap: these words lie past the forward read */
				const uint32_t *rp = rnt + (pos >> 3);
				const int sh = (pos & 7) * 4;
				unsigned lo = rp[0];
				for (; k < nwords; k++) {
					const unsigned hi = *++rp;
					unsigned nt = __funnelshift_r(lo, hi, sh);
					lo = hi;
					if (k == nwords - 1)
						nt &= pb::nibmask(seq_len - 8 * k);
					emit(k, nt);
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
						else if (kind == PB_FILTER_PEAR_TEST)
							pass = pear_test_pass(prm->pear_cdf, prm->filters[k].dvalue, prm->filters[k].dvalue2, prm->filters[k].dvalue3,
							                      bestov, mism, F, R);
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
		if (pair >= 0 && status != ST_DEFER) {
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
				defer_list[dbase + __popc(m_defer & pb::lanemask_lt())] = pair + pair_base;      /* the list is the whole batch's: a slice's kernel counts from the slice */
		}
		__syncwarp();      /* every lane is done with its record before the next batch lands on it */
		batch = batch1;
		batch1 = batch2;
		batch2 = batch3;
		cur = nxt;
		nxt = far;
		pair2 = pair3;
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
	constexpr size_t HEAD = ((LUT_DOUBLES * sizeof(double) + PB_NCOUNTERS * sizeof(unsigned)) + 127) & ~(size_t) 127;
	return HEAD + sizeof(LaneArea<ML>) * WARPS;
}

}  // namespace pbl
