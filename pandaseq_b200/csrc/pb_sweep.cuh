/* pb_sweep.cuh -- K1-K3 of align() (assembler.c:84-118) as a bit-parallel sweep over diagonals, one LANE per read pair.
 *
 * What the reference computes with its 65536 x 2 position table is a join: overlap o is a candidate iff some valid
 * forward 8-mer position p and some valid reverse 8-mer position share a 16-bit code and lie on diagonal o, p being
 * one of the first two forward positions with that code ("table-free statement of seeding").
 * pb::seed_kernel does that join with a shared-memory hash table, one warp per pair, 733 warp-instructions per pair:
 * 284 hash operations at 32 per instruction.  Here the join is never formed.  With the reads as two bit planes of
 * 2-bit k-mer digits (32 bases per word), the forward read shifted by s = F - o against the template-order reverse read
 * gives, per 32-bit word, the positions where the digits differ (XOR, OR); an 8-mer match on that diagonal is a run
 * of eight equal positions, found with three shift-OR doubling steps.  All diagonals of a pair cost
 * (NW (NW + 1) / 2) x 32 word steps of nine instructions for reads of up to 32 NW bases, every lane busy with its own
 * pair (NW = 5, 8 or 10: the length classes of pb_device.cu): measured 230 warp-instructions per pair at 2x150 with the planes and
 * the certificate (profiles/r2/), no shared-memory traffic and no atomics in the sweep itself.
 *
 * The sweep finds every diagonal with a shared 8-mer; the reference only those where the forward position was one of
 * the first two with its code ("lost k-mers", assembler.c:95-97).  So every flagged diagonal is certified afterwards:
 * the lowest matching forward position p on it is taken and the occurrences of its code before p are counted (the
 * 8 digits against all forward positions, bit-parallel again).  Fewer than two: p is in the table, the diagonal is a
 * candidate for certain.  Two or more (0.01 % of random pairs, low-complexity reads): the pair is handed to the general
 * kernel, whose hash join is exact.  Pairs with a base that is not A/C/G/T, reads outside 16 .. 32 NW, or no candidate
 * at all are handed on as pb::seed_kernel does.  The output is pb::seed_kernel's: the candidate mask in overlap order,
 * a flag word and the overlap bin (pb_kernels.cuh), so the lane-per-pair kernel (pb_lanes.cuh) takes either.
 *
 * The functions up to sweep_resolve() are plain C++ on 32-bit words and compile for the host as well: tests/c/
 * sweep_host.cpp runs them against the oracle's table-based seeding on the CPU (tests/test_sweep_core.py).
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PBS_HD __host__ __device__ __forceinline__
#define PBS_HDM __host__ __device__ __forceinline__
#else
#define PBS_HD static inline
#define PBS_HDM inline
#endif

namespace pbs {

constexpr uint32_t NIB1 = 0x11111111u;
constexpr unsigned SEED_GENERAL = 1u, SEED_SKIP = 2u;      /* = pb::PB_SEED_GENERAL / PB_SEED_SKIP */

/* low 32 bits of (hi:lo) >> s, s in [0, 31] */
PBS_HD uint32_t shf_r(uint32_t lo, uint32_t hi, unsigned s) {
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, s);
#else
	return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
/* high 32 bits of (hi:lo) << s, s in [0, 31] */
PBS_HD uint32_t shf_l(uint32_t lo, uint32_t hi, unsigned s) {
#if defined(__CUDA_ARCH__)
	return __funnelshift_l(lo, hi, s);
#else
	return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
PBS_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}
PBS_HD int ffs32(uint32_t x) {      /* 1-based, 0 for x == 0 */
#if defined(__CUDA_ARCH__)
	return __ffs((int) x);
#else
	return __builtin_ffs((int) x);
#endif
}
/* the low n bits set, n in [0, 32] */
PBS_HD uint32_t lowbits(int n) {
	return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
}
/* bits 4k of d (k = 0..7) -> 8 contiguous bits: one multiply lands them at 9..12 and 25..28 without carries
 * (the 32 partial products 4i + 3j, i < 8, j < 4, are all at different positions) */
PBS_HD uint32_t gather8(uint32_t d) {
	const uint32_t m = d * 0x249u;
	return ((m >> 9) & 0xFu) | ((m >> 21) & 0xF0u);
}

/* Bit planes of one read: bit b of p0[j] / p1[j] = low / high bit of the k-mer digit (misc.h:41: T = 3, G = 2, C = 1,
 * everything else 0) of base 32 j + b; zero past the read.  nt = the read's packed 4-bit bases, 8 per word.
 * `bad` collects nibbles that are not exactly A, C, G or T. */
template <int NW>
PBS_HD void build_planes(const uint32_t *nt, int len, uint32_t (&p0)[NW], uint32_t (&p1)[NW], uint32_t &bad) {
	const int nwords = (len + 7) >> 3;
	int bits = 0;
	uint32_t two = 0;
#pragma unroll
	for (int j = 0; j < NW; j++) {
		uint32_t a0 = 0, a1 = 0;
#pragma unroll
		for (int t = 0; t < 4; t++) {
			const int k = 4 * j + t;
			const uint32_t w = k < nwords ? nt[k] : 0u;
			const uint32_t w1 = w >> 1, w2 = w >> 2, w3 = w >> 3;
			a0 |= gather8((w1 | w3) & NIB1) << (8 * t);      /* C or T */
			a1 |= gather8((w2 | w3) & NIB1) << (8 * t);      /* G or T */
			/* exactly one bit per base: no nibble with two bits, and as many bits as bases (the padding nibbles are zero) */
			two |= (w & w1 & 0x77777777u) | (w & w2 & 0x33333333u) | (w & w3 & NIB1);
			bits += popc32(w);
		}
		p0[j] = a0;
		p1[j] = a1;
	}
	bad |= two | (uint32_t) (bits ^ len);
}

/* One word step of the sweep: positions of T word k against the forward planes shifted to them.
 * neq bit i = the digits differ, or the forward position is past the read. */
PBS_HD uint32_t neq_word(uint32_t f0, uint32_t f1, uint32_t fv, uint32_t t0, uint32_t t1) {
#if defined(__CUDA_ARCH__)
	/* two three-input logic operations, spelled out: left to itself the compiler forms three */
	uint32_t a, ne;
	asm("lop3.b32 %0, %1, %2, %3, 0x7D;" : "=r"(a) : "r"(f0), "r"(t0), "r"(fv));       /* (f0 ^ t0) | ~fv */
	asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(ne) : "r"(a), "r"(f1), "r"(t1));        /* a | (f1 ^ t1) */
	return ne;
#else
	const uint32_t a = (f0 ^ t0) | ~fv;
	return a | (f1 ^ t1);
#endif
}

/* 2, 4 and 16 as run-time values (kernel parameters): x * m as a 32 x 32 -> 64-bit multiply gives x << s in the low word
 * and the bits that cross into the next word in the high word, in ONE instruction of the FMA pipe (IMAD.WIDE) -- the
 * sweep is bound by the ALU pipe (LOP3 / SHF), which a funnel shift would load further.  As compile-time constants the
 * compiler turns them back into shifts. */
struct Muls { uint32_t m1, m2, m4; };
PBS_HD void shl_wide(uint32_t x, uint32_t m, uint32_t &lo, uint32_t &hi) {
	const unsigned long long w = (unsigned long long) x * m;
	lo = (uint32_t) w;
	hi = (uint32_t) (w >> 32);
}
/* Run detection over the words of one diagonal, lowest word first: step() returns, for the word of differences it is
 * given, bit i CLEAR iff no position among i-7 .. i differs (three shift-OR doubling steps; the windows reach into the
 * word below through the carries).  A k-mer that ends before position 8 does not exist (misc.h:41-45: nine clean bases),
 * so the first word's bits 0..7 are forced. */
struct Run8 {
	uint32_t c1 = 0, c2 = 0, c4 = 0xFFu;
	PBS_HDM uint32_t step(uint32_t ne, const Muls &mu) {
		uint32_t lo, hi;
		shl_wide(ne, mu.m1, lo, hi);
		const uint32_t y1 = ne | lo | c1;
		c1 = hi;
		shl_wide(y1, mu.m2, lo, hi);
		const uint32_t y2 = y1 | lo | c2;
		c2 = hi;
		shl_wide(y2, mu.m4, lo, hi);
		const uint32_t y4 = y2 | lo | c4;
		c4 = hi;
		return y4;
	}
};

/* Diagonals s = 32 W + b for b in [b0, b1) (overlap o = F - s): mask[W] bit b set iff the reads share a valid 8-mer on
 * it, i.e. eight equal digits at template positions i-7 .. i with 8 <= i < o.  TOP = false leaves out the highest word
 * of every diagonal (and with it the diagonals of the last block, which have only that word): right when no pair of the
 * warp has an overlap that reaches it, i.e. b >= Fmax - 32 (NW - 1). */
template <int NW, bool TOP>
PBS_HD void sweep_range(int b0, int b1, uint32_t (&f0)[NW], uint32_t (&f1)[NW], uint32_t (&fv)[NW],
                        const uint32_t (&t0)[NW], const uint32_t (&t1)[NW], uint32_t (&mask)[NW], uint32_t &bitb, const Muls mu) {
	constexpr int NK = TOP ? NW : NW - 1;      /* words of the longest diagonal */
#pragma unroll 1
	for (int b = b0; b < b1; b++) {
#pragma unroll
		for (int W = 0; W < NK; W++) {
			uint32_t all = 0xFFFFFFFFu;      /* AND of the words' windows: all ones <=> no match on this diagonal */
			Run8 run;
#pragma unroll
			for (int k = 0; k + W < NK; k += 2) {
				const uint32_t ya = run.step(neq_word(f0[k + W], f1[k + W], fv[k + W], t0[k], t1[k]), mu);
				const uint32_t yb = k + 1 + W < NK ? run.step(neq_word(f0[k + 1 + W], f1[k + 1 + W], fv[k + 1 + W], t0[k + 1], t1[k + 1]), mu) : 0xFFFFFFFFu;
				all &= ya & yb;
			}
			mask[W] += (all != 0xFFFFFFFFu ? 1u : 0u) * bitb;
		}
		bitb *= mu.m1;
		/* the forward planes one position further */
#pragma unroll
		for (int j = 0; j < NW; j++) {
			const bool last = j + 1 == NW;
			f0[j] = shf_r(f0[j], last ? 0u : f0[j + 1], 1);
			f1[j] = shf_r(f1[j], last ? 0u : f1[j + 1], 1);
			fv[j] = shf_r(fv[j], last ? 0u : fv[j + 1], 1);
		}
	}
}
/* All diagonals.  fmax = the longest forward read among the pairs swept together (the warp's 32 pairs; the pair's own F
 * on the host): a template position 32 (NW - 1) or higher lies inside an overlap only while b < fmax - 32 (NW - 1).
 * f0 / f1 / fv (fv = positions inside the forward read) are consumed. */
template <int NW>
PBS_HD void sweep(uint32_t (&f0)[NW], uint32_t (&f1)[NW], uint32_t (&fv)[NW],
                  const uint32_t (&t0)[NW], const uint32_t (&t1)[NW], uint32_t (&mask)[NW], int fmax, const Muls mu) {
#pragma unroll
	for (int W = 0; W < NW; W++)
		mask[W] = 0;
	uint32_t bitb = 1;
	int bcut = fmax - 32 * (NW - 1);
	bcut = bcut < 0 ? 0 : (bcut > 32 ? 32 : bcut);
	sweep_range<NW, true>(0, bcut, f0, f1, fv, t0, t1, mask, bitb, mu);
	sweep_range<NW, false>(bcut, 32, f0, f1, fv, t0, t1, mask, bitb, mu);
}

/* Where the certificate can index a pair's words with run-time positions: PL(w) = word w of this lane, PL.set(w, v).
 * Layout: F0[NW + 1], F1[NW + 1] (one zero word on top), T0[NW], T1[NW], the sweep's mask[NW], the candidate mask CW[NW],
 * the postponed counts PEND[NW]. */
template <int NW> struct PlaneIndex {
	static constexpr int F0 = 0, F1 = NW + 1, T0 = 2 * (NW + 1), T1 = T0 + NW, MASK = T1 + NW, CW = MASK + NW, PEND = CW + NW, NPEND = NW,
	                     WORDS = PEND + NPEND;      /* PEND: the certificate's postponed counts (position p | candidate bit << 16) */
};

/* Lowest template position i with an 8-mer match ending there on diagonal s (o = F - s), or -1.  Same arithmetic as
 * sweep(), with positions taken at run time. */
template <int NW, typename PL>
PBS_HD int first_hit(const PL &pl, int s, int o, const Muls &mu) {
	using PI = PlaneIndex<NW>;
	const int w0 = s >> 5, sh = s & 31;
	Run8 run;
	for (int k = 0; 32 * k < o && k + w0 < NW; k++) {
		const uint32_t f0 = shf_r(pl(PI::F0 + w0 + k), pl(PI::F0 + w0 + k + 1), sh);
		const uint32_t f1 = shf_r(pl(PI::F1 + w0 + k), pl(PI::F1 + w0 + k + 1), sh);
		const uint32_t hit = ~run.step(neq_word(f0, f1, lowbits(o - 32 * k), pl(PI::T0 + k), pl(PI::T1 + k)), mu);
		if (hit)
			return 32 * k + ffs32(hit) - 1;
	}
	return -1;
}

/* Forward positions p' in [8, p) whose 8-mer (digits at p'-7 .. p') equals the one ending at p. */
template <int NW, typename PL>
PBS_HD int earlier_occurrences(const PL &pl, int p) {
	using PI = PlaneIndex<NW>;
	const int a = p - 7;
	const uint32_t pat0 = shf_r(pl(PI::F0 + (a >> 5)), pl(PI::F0 + (a >> 5) + 1), a & 31) & 0xFFu;
	const uint32_t pat1 = shf_r(pl(PI::F1 + (a >> 5)), pl(PI::F1 + (a >> 5) + 1), a & 31) & 0xFFu;
	uint32_t m0[8], m1[8], prev[8];
#pragma unroll
	for (int k = 0; k < 8; k++) {
		m0[k] = 0u - ((pat0 >> k) & 1u);
		m1[k] = 0u - ((pat1 >> k) & 1u);
		prev[k] = 0xFFFFFFFFu;                               /* below position 0 nothing matches */
	}
	int count = 0;
	for (int j = 0; 32 * j < p; j++) {
		const uint32_t f0 = pl(PI::F0 + j), f1 = pl(PI::F1 + j);
		uint32_t differs = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) {
			/* digit k of the pattern sits 7 - k positions below the position the window ends at */
			const uint32_t ne = (f0 ^ m0[k]) | (f1 ^ m1[k]);
			differs |= shf_l(prev[k], ne, 7 - k);
			prev[k] = ne;
		}
		uint32_t occ = ~differs & lowbits(p - 32 * j);
		if (j == 0)
			occ &= ~0xFFu;
		count += popc32(occ);
	}
	return count;
}

/* From the sweep's diagonal mask to pb::seed_kernel's record: PL's CW words = candidate mask in overlap order (bit i <=>
 * overlap mo + i, assembler.c:39), `flags` the flag word, `low` the lowest candidate bit (the lane kernel's bin).
 * maxov as assembler.c:78-82 with maxoverlap == 0, i.e. min(F, R); the caller guarantees mo < maxov <= 32 NW.
 *
 * The certificate.  A flagged diagonal s has a match (p, q) = (s + i, i).  The reference misses it only if p is "lost":
 * two valid forward positions p1 < p2 < p carry p's code.  Each of them matches q as well, on diagonal p_x - q < s --
 * flagged too, if it is not negative -- or lies in [8, q).  So with `seen` flagged diagonals below s and the lowest
 * match of this one at i, seen + (i - 8) < 2 proves p is in the table; only otherwise are p's earlier occurrences
 * counted, and that is put off to a second phase so that the lanes of a warp do it together (step1 / step2 are one
 * flagged diagonal / one postponed count per call; the kernel calls them while any lane has work). */
template <int NW, typename PL> struct Resolver {
	using PI = PlaneIndex<NW>;
	PL &pl;
	int F, mo, maxov;
	Muls mu;
	int W = -1, seen = 0, low = 1 << 20, npend = 0;
	uint32_t m = 0;
	unsigned flags = 0;
	PBS_HDM Resolver(PL &pl_, int F_, int mo_, int maxov_, const Muls &mu_) : pl(pl_), F(F_), mo(mo_), maxov(maxov_), mu(mu_) {}
	PBS_HDM void accept(int idx) {
		pl.set(PI::CW + (idx >> 5), pl(PI::CW + (idx >> 5)) | (1u << (idx & 31)));
		low = idx < low ? idx : low;
	}
	PBS_HDM bool step1() {
		while (m == 0) {
			if (++W >= NW)
				return false;
			m = pl(PI::MASK + W);
		}
		const int b = ffs32(m) - 1;
		m &= m - 1;
		const int s = 32 * W + b, o = F - s, idx = o - mo;
		const int below = seen++;
		if (idx < 0 || o > maxov)           /* BIT_LIST_SET (assembler.c:39): outside the bit list */
			return true;
		const int i = first_hit<NW>(pl, s, o, mu);
		if (i < 0)
			return true;                    /* cannot happen: the sweep saw a match on this diagonal */
		if (below + (i - 8) < 2) {
			accept(idx);
		} else if (npend < PI::NPEND) {
			pl.set(PI::PEND + npend, (uint32_t) (s + i) | ((uint32_t) idx << 16));
			npend++;
		} else {
			flags |= SEED_GENERAL;          /* more uncertified diagonals than this keeps: the exact join decides */
		}
		return true;
	}
	PBS_HDM bool step2() {
		if (npend == 0)
			return false;
		npend--;
		const uint32_t req = pl(PI::PEND + npend);
		const int p = (int) (req & 0xFFFFu), idx = (int) (req >> 16);
		if (earlier_occurrences<NW>(pl, p) >= 2)
			flags |= SEED_GENERAL;          /* maybe a lost k-mer (assembler.c:95-97): the exact join decides */
		else
			accept(idx);
		return true;
	}
	PBS_HDM unsigned finish() {
		if (low == (1 << 20))
			flags |= SEED_GENERAL;          /* ALL_BITS_IF_NONE (assembler.c:118): every overlap is scored, the general kernel's job */
		return flags;
	}
};

}  // namespace pbs

#if defined(__CUDACC__)
#include "pb_kernels.cuh"

namespace pbs {

/* Per-warp shared memory: the 32 pairs' packed bases as the bulk copies land them, and the pairs' planes, word-major
 * (word w of lane l at [w][l]) so that a warp reading one word index never collides.  The two share their memory: the bases
 * are dead once every lane has built its planes (a __syncwarp() apart), and the next batch's copies are issued only after
 * every lane is done with its planes. */
template <int NW> struct SweepArea {
	static constexpr int ML = 32 * NW;
	static constexpr int NT_BYTES = ML;                                  /* 2 reads x ML / 2 bytes */
	static constexpr int STRIDE = (((NT_BYTES + 15) / 16) | 1) * 16;     /* odd number of 16-byte units */
	union alignas(128) {
		uint8_t nt[32 * STRIDE];
		uint32_t planes[PlaneIndex<NW>::WORDS][32];
	};
	alignas(8) uint64_t bar;
};

struct LanePlanes {
	uint32_t *base;      /* &planes[0][lane] */
	__device__ __forceinline__ uint32_t operator()(int w) const { return base[w * 32]; }
	__device__ __forceinline__ void set(int w, uint32_t v) const { base[w * 32] = v; }
};

/* seeds / bin_count as pb::seed_kernel writes them (pb_kernels.cuh).  A warp takes 32 consecutive pairs at a time; which 32
 * comes from a counter in global memory (*next_batch, zero at launch): the scheduler favours some warps of an SM over others,
 * and with a fixed share per warp the favoured ones would sit at the final barrier while the rest finish alone. */
template <int NW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
sweep_seed_kernel(const pb_device_params *__restrict__ prm, int n, const uint8_t *__restrict__ reads,
                  const pb_pair_meta *__restrict__ meta, uint32_t *__restrict__ seeds, unsigned *__restrict__ bin_count,
                  unsigned *__restrict__ next_batch, const Muls mu, const int *__restrict__ list, const int *__restrict__ list_n) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	__shared__ unsigned s_bins[pb::PB_SEED_BINS];
	using SA = SweepArea<NW>;
	using PI = PlaneIndex<NW>;
	constexpr int ML = 32 * NW;
	constexpr int MW = pb::seed_mask_words(ML), SWORDS = pb::seed_words(ML);
	static_assert(MW == NW, "one mask word per plane word");
	SA *areas = reinterpret_cast<SA *>(smem_raw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < pb::PB_SEED_BINS)
		s_bins[tid] = 0;
	SA &wa = areas[warp];
	if (lane == 0) {
		pb::mbar_init(&wa.bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const int mo = prm->minoverlap;
	if (list_n)
		n = *list_n;          /* the pairs of one length class (pb::class_list_kernel), listed by index */
	const int nbatch = (n + 31) >> 5;
	unsigned parity = 0;
	uint32_t *const myplanes = &wa.planes[0][lane];
	LanePlanes pl{myplanes};

	for (;;) {
		int batch = 0;
		if (lane == 0)
			batch = (int) atomicAdd(next_batch, 1u);
		batch = __shfl_sync(pb::FULL, batch, 0);
		if (batch >= nbatch)
			break;
		const int item = batch * 32 + lane;
		const int pair = item < n ? (list ? list[item] : item) : -1;
		unsigned off16 = 0;
		int F = 0xFFFF, R = 0;
		if (pair >= 0) {
			const uint2 mraw = *reinterpret_cast<const uint2 *>(&meta[pair]);
			off16 = mraw.x;
			F = (int) (mraw.y & 0xFFFFu);
			R = (int) (mraw.y >> 16);
		}
		unsigned flags = 0;
		if (F == 0xFFFF)
			flags = SEED_SKIP;
		else if (F > ML || R > ML || F < 16 || R < 16 || mo >= min(F, R))
			flags = SEED_GENERAL;
		const int fw = (F + 7) >> 3, rw = (R + 7) >> 3;
		const unsigned bytes = flags ? 0u : (unsigned) ((fw + rw) * 4 + 15) & ~15u;
		const unsigned total = __reduce_add_sync(pb::FULL, bytes);
		/* the copies below (async proxy) land where this warp's lanes wrote their planes of the previous batch (generic proxy) */
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if (lane == 0)
			pb::mbar_expect_tx(&wa.bar, total);
		__syncwarp();
		if (bytes)
			pb::bulk_g2s(wa.nt + lane * SA::STRIDE, reads + (size_t) off16 * 16, bytes, &wa.bar);
		pb::mbar_wait(&wa.bar, parity);
		parity ^= 1u;

		int lowest = 1 << 20;
		{
			uint32_t f0[NW], f1[NW], fv[NW], t0[NW], t1[NW], mask[NW];
			uint32_t bad = 0;
			const uint32_t *nt = reinterpret_cast<const uint32_t *>(wa.nt + lane * SA::STRIDE);
			const int Fe = flags ? 0 : F, Re = flags ? 0 : R;          /* lanes without a pair sweep empty reads */
			build_planes<NW>(nt, Fe, f0, f1, bad);
			build_planes<NW>(nt + fw, Re, t0, t1, bad);
			if (bad)
				flags |= SEED_GENERAL;
			__syncwarp();      /* the planes go where the bases were: every lane has read its own */
#pragma unroll
			for (int j = 0; j < NW; j++) {
				fv[j] = lowbits(min(max(Fe - 32 * j, 0), 32));
				myplanes[(PI::F0 + j) * 32] = f0[j];
				myplanes[(PI::F1 + j) * 32] = f1[j];
				myplanes[(PI::T0 + j) * 32] = t0[j];
				myplanes[(PI::T1 + j) * 32] = t1[j];
			}
			myplanes[(PI::F0 + NW) * 32] = 0;
			myplanes[(PI::F1 + NW) * 32] = 0;
			/* the longest forward read of the warp bounds the template positions any overlap reaches (warp-uniform) */
			const int fmax = (int) __reduce_max_sync(pb::FULL, (unsigned) Fe);
			sweep<NW>(f0, f1, fv, t0, t1, mask, fmax, mu);
#pragma unroll
			for (int j = 0; j < NW; j++) {
				myplanes[(PI::MASK + j) * 32] = flags ? 0u : mask[j];
				myplanes[(PI::CW + j) * 32] = 0;
			}
		}
		{
			Resolver<NW, LanePlanes> res(pl, F, mo, min(F, R), mu);
			bool more = flags == 0;
			while (__any_sync(pb::FULL, more)) {
				if (more)
					more = res.step1();
			}
			more = flags == 0;
			while (__any_sync(pb::FULL, more)) {
				if (more)
					more = res.step2();
			}
			if (!flags) {
				flags = res.finish();
				lowest = res.low;
			}
		}
		unsigned bin = flags ? (unsigned) (pb::PB_SEED_BINS - 1) : (unsigned) (lowest >> 4);
		if (bin > (unsigned) (pb::PB_SEED_BINS - 1))
			bin = pb::PB_SEED_BINS - 1;
		if (pair >= 0) {
			uint32_t sw[SWORDS];
#pragma unroll
			for (int w = 0; w < SWORDS; w++)
				sw[w] = w < MW ? myplanes[(PI::CW + w) * 32] : 0u;
			sw[MW] = flags;
			sw[MW + 1] = bin;
			uint4 *dst = reinterpret_cast<uint4 *>(seeds + (size_t) pair * SWORDS);
#pragma unroll
			for (int q = 0; q < SWORDS / 4; q++)
				dst[q] = make_uint4(sw[4 * q], sw[4 * q + 1], sw[4 * q + 2], sw[4 * q + 3]);
			atomicAdd(&s_bins[bin], 1u);
		}
		__syncwarp();      /* every lane is done with its slot before the next batch lands on it */
	}
	__syncthreads();
	if (tid < pb::PB_SEED_BINS && s_bins[tid])
		atomicAdd(&bin_count[tid], s_bins[tid]);
}

template <int NW, int WARPS> constexpr size_t sweep_smem_bytes() {
	return sizeof(SweepArea<NW>) * WARPS;
}

}  // namespace pbs
#endif
