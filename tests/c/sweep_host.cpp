/* sweep_host.cpp -- TEST HARNESS: the host-compilable core of pandaseq_b200/csrc/pb_sweep.cuh (plane building, the diagonal
 * sweep, the certificate) run on the CPU over flat panda_qual pairs, so that tests/test_sweep_core.py can compare the
 * candidate masks with the oracle's table-based seeding (oracle/panda_oracle.c po_seed_bits) without a GPU.
 * Compiled by the test with g++; not part of the product library. */
#include <stdint.h>
#include <string.h>
#include <vector>
#include "pb_sweep.cuh"

template <int NW> struct HostPlanes {
	uint32_t w[pbs::PlaneIndex<NW>::WORDS];
	uint32_t operator()(int i) const { return w[i]; }
	void set(int i, uint32_t v) { w[i] = v; }
};

/* one pair: pack as pb::pack_kernel does (4-bit codes, reverse read in template order), then the kernel's steps */
template <int NW>
static unsigned one_pair(const uint8_t *f, int F, const uint8_t *r, int R, int mo, int fmax, uint32_t *cw_out, int *lowest) {
	using PI = pbs::PlaneIndex<NW>;
	constexpr int ML = 32 * NW;
	for (int w = 0; w < NW; w++)
		cw_out[w] = 0;
	*lowest = 1 << 20;
	if (F > ML || R > ML || F < 16 || R < 16 || mo >= (F < R ? F : R))
		return pbs::SEED_GENERAL;
	uint32_t fnt[ML / 8 + 1] = { 0 }, rnt[ML / 8 + 1] = { 0 };
	for (int i = 0; i < F; i++)
		fnt[i >> 3] |= (uint32_t) (f[2 * i] & 15u) << (4 * (i & 7));
	for (int i = 0; i < R; i++)
		rnt[i >> 3] |= (uint32_t) (r[2 * (R - 1 - i)] & 15u) << (4 * (i & 7));
	uint32_t f0[NW], f1[NW], fv[NW], t0[NW], t1[NW], mask[NW], bad = 0;
	const pbs::Muls mu = { 2u, 4u, 16u };
	pbs::build_planes<NW>(fnt, F, f0, f1, bad);
	pbs::build_planes<NW>(rnt, R, t0, t1, bad);
	HostPlanes<NW> pl;
	memset(&pl, 0, sizeof pl);
	for (int j = 0; j < NW; j++) {
		int nb = F - 32 * j;
		fv[j] = pbs::lowbits(nb < 0 ? 0 : (nb > 32 ? 32 : nb));
		pl.w[PI::F0 + j] = f0[j];
		pl.w[PI::F1 + j] = f1[j];
		pl.w[PI::T0 + j] = t0[j];
		pl.w[PI::T1 + j] = t1[j];
	}
	if (bad)
		return pbs::SEED_GENERAL;
	pbs::sweep<NW>(f0, f1, fv, t0, t1, mask, fmax > F ? fmax : F, mu);
	for (int j = 0; j < NW; j++)
		pl.w[PI::MASK + j] = mask[j];
	pbs::Resolver<NW, HostPlanes<NW>> res(pl, F, mo, F < R ? F : R, mu);
	while (res.step1()) {}
	while (res.step2()) {}
	unsigned flags = res.finish();
	*lowest = res.low;
	for (int w = 0; w < NW; w++)
		cw_out[w] = pl.w[PI::CW + w];
	return flags;
}

/* f_data / r_data: {nt, qual} byte pairs; offsets in elements.  cw: n x 16 words, flags / lowest: n. */
extern "C" int sweep_host_run(int nw, size_t n, const uint8_t *f_data, const uint64_t *f_off, const uint8_t *r_data, const uint64_t *r_off,
                              int minoverlap, uint32_t *cw, uint32_t *flags, int32_t *lowest) {
	for (size_t i = 0; i < n; i++) {
		/* the kernel sweeps 32 pairs together and leaves out the top word where the longest forward read of the 32 allows */
		int fmax = 0;
		for (size_t j = i & ~(size_t) 31; j < n && j < (i | 31) + 1; j++) {
			const int Fj = (int) (f_off[j + 1] - f_off[j]);
			if (Fj <= 32 * nw && Fj > fmax)
				fmax = Fj;
		}
		const uint8_t *f = f_data + 2 * f_off[i], *r = r_data + 2 * r_off[i];
		const int F = (int) (f_off[i + 1] - f_off[i]), R = (int) (r_off[i + 1] - r_off[i]);
		uint32_t *c = cw + 16 * i;
		memset(c, 0, 16 * sizeof(uint32_t));
		int low = 0;
		switch (nw) {
		case 3: flags[i] = one_pair<3>(f, F, r, R, minoverlap, fmax, c, &low); break;
		case 5: flags[i] = one_pair<5>(f, F, r, R, minoverlap, fmax, c, &low); break;
		case 8: flags[i] = one_pair<8>(f, F, r, R, minoverlap, fmax, c, &low); break;
		case 10: flags[i] = one_pair<10>(f, F, r, R, minoverlap, fmax, c, &low); break;
		default: return -1;
		}
		lowest[i] = low;
	}
	return 0;
}
