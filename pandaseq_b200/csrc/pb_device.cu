/* pb_device.cu -- device context, launch wrappers and the host-buffer (e2e) path.
 *
 * extern "C" entry points declared in include/pandaseq_b200.h.  One pb_context
 * per (process, GPU): a stream, the parameter/LUT block in HBM, and staging
 * buffers for the host-buffer path.  Nothing here computes on reads on the CPU:
 * the host only lays out offsets and moves bytes.
 */
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pb_kernels.cuh"
#include "pb_lanes.cuh"
#include "pb_sweep.cuh"

#include "pb_ctx.h"

extern "C" int pb_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess)
		return 0;
	return n;
}

extern "C" pb_status pb_context_create(int device, pb_context **out) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		pb_set_error("no CUDA device available (%s); libpandaseq_b200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
		return PB_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= n) {
		pb_set_error("device %d out of range (0..%d)", device, n - 1);
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(device));
	pb_context *ctx = (pb_context *) calloc(1, sizeof(pb_context));
	if (!ctx)
		return PB_ERR_NOMEM;
	ctx->device = device;
	ctx->lanes_mode = -1;
	pthread_mutex_init(&ctx->lock, NULL);
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	ctx->sm_count = prop.multiProcessorCount;
	if (prop.major < 10) {
		pb_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
		free(ctx);
		return PB_ERR_NO_DEVICE;
	}
	CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaMalloc(&ctx->d_params, sizeof(pb_device_params)));
	CUDA_TRY(cudaMallocHost(&ctx->h_params, sizeof(pb_device_params)));
	CUDA_TRY(cudaMalloc(&ctx->d_counters, PB_NCOUNTERS * sizeof(unsigned long long)));
	for (int k = 0; k < 4; k++)
		CUDA_TRY(cudaEventCreate(&ctx->tev[k]));
	CUDA_TRY(cudaMalloc(&ctx->d_defer_total, sizeof(unsigned long long)));
	CUDA_TRY(cudaMemset(ctx->d_defer_total, 0, sizeof(unsigned long long)));
	for (int s = 0; s < PB_HOST_SLOTS; s++)
		CUDA_TRY(cudaEventCreateWithFlags(&ctx->slot[s].done, cudaEventDisableTiming));
	*out = ctx;
	return PB_OK;
}

static void free_slot(pb_context::Slot &s) {
	cudaFreeHost(s.h_f); cudaFreeHost(s.h_r); cudaFreeHost(s.h_foff); cudaFreeHost(s.h_roff); cudaFreeHost(s.h_recoff);
	cudaFree(s.d_f); cudaFree(s.d_r); cudaFree(s.d_foff); cudaFree(s.d_roff); cudaFree(s.d_recoff);
	cudaFree(s.d_reads); cudaFree(s.d_meta); cudaFree(s.d_res); cudaFreeHost(s.h_res);
	cudaFree(s.d_nt); cudaFreeHost(s.h_nt); cudaFree(s.d_p); cudaFreeHost(s.h_p);
	cudaFree(s.d_code); cudaFreeHost(s.h_code); cudaFreeHost(s.h_meta); cudaFreeHost(s.h_reads);
	cudaEvent_t ev = s.done;
	memset(&s, 0, sizeof s);
	s.done = ev;
}

extern "C" void pb_context_destroy(pb_context *ctx) {
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	cudaStreamSynchronize(ctx->copy_stream);
	for (int s = 0; s < PB_HOST_SLOTS; s++) {
		free_slot(ctx->slot[s]);
		cudaEventDestroy(ctx->slot[s].done);
	}
	cudaFree(ctx->d_params);
	cudaFreeHost(ctx->h_params);
	cudaFree(ctx->d_counters);
	cudaFree(ctx->d_scratch);
	cudaFree(ctx->d_defer[0]);
	cudaFree(ctx->d_defer[1]);
	cudaFree(ctx->d_seeds[0]);
	cudaFree(ctx->d_seeds[1]);
	cudaFree(ctx->d_order[0]);
	cudaFree(ctx->d_order[1]);
	if (ctx->ovl_ready) {
		cudaStreamSynchronize(ctx->ovl_stream);
		for (int k = 0; k < 18; k++)
			cudaEventDestroy(ctx->ovl_ev[k]);
		cudaStreamDestroy(ctx->ovl_stream);
		cudaFree(ctx->d_ovl);
	}
	cudaFree(ctx->d_classes[0]);
	cudaFree(ctx->d_classes[1]);
	cudaFree(ctx->d_bins[0]);
	cudaFree(ctx->d_bins[1]);
	cudaFree(ctx->d_defer_total);
	cudaFree(ctx->d_pear_cdf);
	for (int k = 0; k < 4; k++)
		cudaEventDestroy(ctx->tev[k]);
	pb_io_release(ctx);
	cudaStreamDestroy(ctx->stream);
	cudaStreamDestroy(ctx->copy_stream);
	free(ctx);
}

extern "C" void *pb_context_stream(pb_context *ctx) {
	return (void *) ctx->stream;
}

extern "C" pb_status pb_synchronize(pb_context *ctx) {
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	return PB_OK;
}

pb_status pb_upload_params(pb_context *ctx, const pb_config *cfg) {
	if (ctx->cfg_valid && memcmp(&ctx->cached_cfg, cfg, sizeof *cfg) == 0)
		return PB_OK;
	/* the pinned mirror may still be in flight from the previous upload */
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	pb_status st = pb_build_device_params(cfg, ctx->h_params);
	if (st != PB_OK)
		return st;
	for (int k = 0; k < cfg->nfilters && k < PB_MAX_FILTERS; k++) {
		if (cfg->filters[k].kind != PB_FILTER_PEAR_TEST)
			continue;
		if (!ctx->d_pear_cdf) {          /* 1.6 MB, once per context */
			const size_t bytes = (size_t) PB_PEAR_ROWS * PB_PEAR_COLS * sizeof(double);
			double *h = (double *) malloc(bytes);
			if (!h)
				return PB_ERR_NOMEM;
			pb_build_pear_cdf(h);
			CUDA_TRY(cudaMalloc(&ctx->d_pear_cdf, bytes));
			CUDA_TRY(cudaMemcpy(ctx->d_pear_cdf, h, bytes, cudaMemcpyHostToDevice));
			free(h);
		}
		ctx->h_params->pear_cdf = ctx->d_pear_cdf;
	}
	CUDA_TRY(cudaMemcpyAsync(ctx->d_params, ctx->h_params, sizeof(pb_device_params), cudaMemcpyHostToDevice, ctx->stream));
	ctx->cached_cfg = *cfg;
	ctx->cfg_valid = true;
	return PB_OK;
}

template <int ML, bool OVER, int WARPS, bool FULLF>
static pb_status launch_assemble(pb_context *ctx, int n, const uint8_t *d_reads, const pb_pair_meta *d_meta,
                                 pb_pair_result *d_results, uint8_t *d_seq_nt, double *d_seq_p, size_t seq_stride,
                                 unsigned long long *d_counters, cudaStream_t stream, bool post,
                                 const int *d_list = nullptr, const int *d_list_n = nullptr, uint16_t *d_seq_code = nullptr) {
	auto kern = pb::assemble_kernel<ML, OVER, WARPS, FULLF>;
	constexpr size_t smem = pb::assemble_smem_bytes<ML, OVER, WARPS>();
	static bool configured[16] = { false };
	if (!configured[ctx->device & 15]) {
		CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured[ctx->device & 15] = true;
	}
	int per_sm = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
	if (per_sm < 1) {
		pb_set_error("assemble kernel does not fit on an SM (smem %zu)", smem);
		return PB_ERR_CUDA;
	}
	/* persistent grid: a whole number of CTAs per SM; warps stride over the batch */
	long long want = ((long long) n + WARPS - 1) / WARPS;
	long long grid = (long long) ctx->sm_count * per_sm;
	if (grid > want)
		grid = want;
	if (grid < 1)
		grid = 1;
	uint8_t *scratch = nullptr;
	if (post) {                  /* primers-after: every resident warp stages its assembled sequence in global scratch */
		const size_t need = (size_t) grid * WARPS * PB_SCRATCH_STRIDE * 2;      /* x2: the host path runs two streams */
		if (need > ctx->scratch_bytes) {
			CUDA_TRY(cudaDeviceSynchronize());
			cudaFree(ctx->d_scratch);
			ctx->d_scratch = nullptr;
			ctx->scratch_bytes = 0;
			CUDA_TRY(cudaMalloc(&ctx->d_scratch, need));
			ctx->scratch_bytes = need;
		}
		/* the halves are fixed by the allocation, not by this launch: the other stream may be running a kernel of another
		 * length class (a smaller grid), and its half must not move under it */
		scratch = ctx->d_scratch + (stream == ctx->copy_stream ? ctx->scratch_bytes / 2 : 0);
	}
	const bool timed = ctx->timing && stream == ctx->stream;
	if (timed && !d_list)
		CUDA_TRY(cudaEventRecord(ctx->tev[0], stream));
	kern<<<(unsigned) grid, WARPS * 32, smem, stream>>>(ctx->d_params, n, d_reads, d_meta, d_results, d_seq_nt, d_seq_p,
	                                                           (long long) seq_stride, d_counters, scratch, d_list, d_list_n, d_seq_code);
	CUDA_TRY(cudaGetLastError());
	if (timed) {
		CUDA_TRY(cudaEventRecord(ctx->tev[3], stream));
		ctx->timing_kind = d_list ? (ctx->timing_overlap ? 3 : 2) : 1;
	}
	return PB_OK;
}

/* The two-kernel path for the common configurations: the seeding kernel (pbs::sweep_seed_kernel, one lane per pair; or
 * pb::seed_kernel, the hash join, one warp per pair) leaves the candidate overlaps of every pair, pbl::assemble_lanes_kernel
 * (lane per pair, K4-K6) scores and merges, and the general kernel assembles the pairs those two handed on.  A batch whose
 * reads are all <= 160 nt runs as one class in batch order; longer or mixed batches are first listed by length class
 * (pb::class_list_kernel) and every class runs the kernels sized for it. */
constexpr int PB_BINS_STRIDE = 2 * pb::PB_SEED_BINS + 4;      /* pairs per bin, cursors, the two kernels' batch counters */
struct LanesState {
	pb_context *ctx;
	int n, si;
	const uint8_t *d_reads;
	const pb_pair_meta *d_meta;
	pb_pair_result *d_results;
	uint8_t *d_seq_nt;
	size_t seq_stride;
	unsigned long long *d_counters;
	cudaStream_t stream;
	int *d_count, *d_list;          /* the general kernel's list: count, entries */
	uint32_t *d_seeds;              /* seeds records (see seeds()) */
	size_t cap;
	uint32_t *seeds(int c) const { return c == 0 && ctx->classes_on[si] ? d_seeds + cap * pb::seed_words(320) : d_seeds; }
	int *class_list(int c) const { return ctx->d_classes[si] + (size_t) c * cap; }
	int *class_count(int c) const { return ctx->d_defer[si] + 1 + c; }
	int *order(int c) const { return ctx->d_order[si] + (size_t) c * cap; }
	unsigned *bins(int c) const { return ctx->d_bins[si] + (size_t) c * PB_BINS_STRIDE; }
};

static pb_status lanes_prepare(LanesState &L, bool classes) {
	pb_context *ctx = L.ctx;
	const int si = L.si;
	const size_t n = (size_t) L.n;
	const bool grow = n + 4 > ctx->defer_cap[si], need_lists = classes && !ctx->classes_on[si];
	if (grow || need_lists) {
		const size_t cap = grow ? n + n / 4 + 64 : ctx->defer_cap[si];
		const bool lists = classes || ctx->classes_on[si];
		CUDA_TRY(cudaDeviceSynchronize());
		cudaFree(ctx->d_defer[si]);
		cudaFree(ctx->d_seeds[si]);
		cudaFree(ctx->d_order[si]);
		cudaFree(ctx->d_classes[si]);
		ctx->d_defer[si] = nullptr;
		ctx->d_seeds[si] = nullptr;
		ctx->d_order[si] = nullptr;
		ctx->d_classes[si] = nullptr;
		ctx->defer_cap[si] = 0;
		ctx->classes_on[si] = false;
		CUDA_TRY(cudaMalloc(&ctx->d_defer[si], (cap + 8) * sizeof(int)));
		/* sized for the widest record; by length class, the 160-nt class (8-word records) has its own region behind the one the two
		 * longer classes (12-word records) share: every class indexes by pair, and all classes are seeded before any is assembled */
		CUDA_TRY(cudaMalloc(&ctx->d_seeds[si], cap * (pb::seed_words(320) + (lists ? pb::seed_words(160) : 0)) * sizeof(uint32_t)));
		CUDA_TRY(cudaMalloc(&ctx->d_order[si], (lists ? pb::PB_LEN_CLASSES : 1) * cap * sizeof(int)));
		if (lists)
			CUDA_TRY(cudaMalloc(&ctx->d_classes[si], pb::PB_LEN_CLASSES * cap * sizeof(int)));
		if (!ctx->d_bins[si])
			CUDA_TRY(cudaMalloc(&ctx->d_bins[si], pb::PB_LEN_CLASSES * PB_BINS_STRIDE * sizeof(unsigned)));      /* bins + the kernels' batch counters, per class */
		ctx->classes_on[si] = lists;
		ctx->defer_cap[si] = cap;
	}
	L.cap = ctx->defer_cap[si];
	L.d_count = ctx->d_defer[si];          /* [0] the general kernel's list length, [1 ..] the class lists' lengths, [8 ..] the general kernel's list */
	L.d_list = ctx->d_defer[si] + 8;
	L.d_seeds = ctx->d_seeds[si];
	CUDA_TRY(cudaMemsetAsync(L.d_count, 0, 8 * sizeof(int), L.stream));
	CUDA_TRY(cudaMemsetAsync(ctx->d_bins[si], 0, pb::PB_LEN_CLASSES * PB_BINS_STRIDE * sizeof(unsigned), L.stream));
	return PB_OK;
}

/* seeding + bin list of one class (c < 0: the whole batch in batch order) */
template <int ML, int SW, int XW>
static pb_status lanes_seed(const LanesState &L, int c, bool sweep) {
	pb_context *ctx = L.ctx;
	const int cc = c < 0 ? 0 : c;
	const int *list = c < 0 ? nullptr : L.class_list(c), *list_n = c < 0 ? nullptr : L.class_count(c);
	const int n = L.n;
	static bool configured[16] = { false }, configured_join[16] = { false };
	if (sweep) {
		auto sweepk = pbs::sweep_seed_kernel<ML / 32, XW>;      /* the diagonal sweep, one lane per pair (pb_sweep.cuh) */
		constexpr size_t sweep_smem = pbs::sweep_smem_bytes<ML / 32, XW>();
		static_assert(ML % 32 == 0 && sweep_smem <= 227 * 1024, "per-CTA shared memory");
		if (!configured[ctx->device & 15]) {
			CUDA_TRY(cudaFuncSetAttribute(sweepk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sweep_smem));
			configured[ctx->device & 15] = true;
		}
		long long grid = (((long long) n + 31) / 32 + XW - 1) / XW;
		if (grid > ctx->sm_count)
			grid = ctx->sm_count;
		const pbs::Muls mu = { 2u, 4u, 16u };      /* run-time values on purpose: pb_sweep.cuh */
		sweepk<<<(unsigned) (grid < 1 ? 1 : grid), XW * 32, sweep_smem, L.stream>>>(ctx->d_params, n, L.d_reads, L.d_meta, L.seeds(c), L.bins(cc),
		                                                                          L.bins(cc) + 2 * pb::PB_SEED_BINS, mu, list, list_n);
	} else {
		if constexpr (SW > 0) {
			auto seedk = pb::seed_kernel<ML, SW>;               /* the hash join, one warp per pair: whole batches only */
			constexpr size_t seed_smem = sizeof(pb::WarpSmem<ML>) * SW;
			static_assert(seed_smem <= 227 * 1024, "per-CTA shared memory");
			if (!configured_join[ctx->device & 15]) {
				CUDA_TRY(cudaFuncSetAttribute(seedk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) seed_smem));
				configured_join[ctx->device & 15] = true;
			}
			long long grid = ((long long) n + SW - 1) / SW;
			if (grid > ctx->sm_count)
				grid = ctx->sm_count;
			seedk<<<(unsigned) (grid < 1 ? 1 : grid), SW * 32, seed_smem, L.stream>>>(ctx->d_params, n, L.d_reads, L.d_meta, L.seeds(c), L.bins(cc));
		} else {
			pb_set_error("internal: no hash-join seeding for this length class");
			return PB_ERR_ARGUMENT;
		}
	}
	CUDA_TRY(cudaGetLastError());
	pb::bin_order_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, L.stream>>>(n, L.seeds(c), pb::seed_words(ML), pb::seed_mask_words(ML) + 1, L.bins(cc), L.order(cc),
	                                                                         list, list_n);
	CUDA_TRY(cudaGetLastError());
	return PB_OK;
}

/* the lane-per-pair kernel over one class's bin list */
template <int LML, int LW>
static pb_status lanes_assemble(const LanesState &L, int c) {
	pb_context *ctx = L.ctx;
	const int cc = c < 0 ? 0 : c;
	auto kern = pbl::assemble_lanes_kernel<LML, LW>;
	constexpr size_t smem = pbl::lanes_smem_bytes<LML, LW>();
	static_assert(smem <= 227 * 1024, "per-CTA shared memory");
	static bool configured[16] = { false };
	if (!configured[ctx->device & 15]) {
		CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured[ctx->device & 15] = true;
	}
	const long long nbatch = ((long long) L.n + 31) / 32;
	long long grid = ((long long) nbatch + LW - 1) / LW;
	if (grid > ctx->sm_count)
		grid = ctx->sm_count;
	if (grid < 1)
		grid = 1;
	kern<<<(unsigned) grid, LW * 32, smem, L.stream>>>(ctx->d_params, L.n, L.d_reads, L.d_meta, L.seeds(c), L.order(cc), L.d_results, L.d_seq_nt, (long long) L.seq_stride,
	                                                     L.d_counters, L.d_list, L.d_count, ctx->d_defer_total, L.bins(cc) + 2 * pb::PB_SEED_BINS + 1,
	                                                     c < 0 ? nullptr : L.class_count(c), 0);
	CUDA_TRY(cudaGetLastError());
	return PB_OK;
}

#define PB_TRY(x) do { pb_status st__ = (x); if (st__ != PB_OK) return st__; } while (0)

/* Large batches of the 152-nt class: the two kernels side by side.  The sweep keeps the ALU pipe busy and little else; the lane kernel
 * waits on dependent scalar code at the few warps its shared memory allows.  Cut into slices, the sweep of slice i + 1 (on the caller's
 * stream) runs next to the lane kernel of slice i (on a second stream), each with fewer warps than alone -- XWO + LWO warps and their
 * shared memory fit one SM together (the sweep's bases and planes share their memory for that) -- and each slice is a batch of its own:
 * pointers offset, pair indices counted from the slice, only the general kernel's list is the whole batch's.  The first sweep and the
 * last lane kernel have the SMs to themselves and run at full width. */
constexpr int PB_OVL_SLICES = 16;
static pb_status ensure_overlap(pb_context *ctx) {
	if (ctx->ovl_ready)
		return PB_OK;
	CUDA_TRY(cudaStreamCreateWithFlags(&ctx->ovl_stream, cudaStreamNonBlocking));
	for (int k = 0; k < PB_OVL_SLICES + 2; k++)
		CUDA_TRY(cudaEventCreateWithFlags(&ctx->ovl_ev[k], cudaEventDisableTiming));
	CUDA_TRY(cudaMalloc(&ctx->d_ovl, PB_OVL_SLICES * PB_BINS_STRIDE * sizeof(unsigned)));
	ctx->ovl_ready = true;
	return PB_OK;
}

template <int NW, int XW>
static pb_status overlap_sweep(pb_context *ctx, int cnt, const uint8_t *d_reads, const pb_pair_meta *meta, uint32_t *seeds, unsigned *bins, cudaStream_t stream) {
	auto k = pbs::sweep_seed_kernel<NW, XW>;
	constexpr size_t smem = pbs::sweep_smem_bytes<NW, XW>();
	static bool configured[16] = { false };
	if (!configured[ctx->device & 15]) {
		CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
		configured[ctx->device & 15] = true;
	}
	long long grid = (((long long) cnt + 31) / 32 + XW - 1) / XW;
	if (grid > ctx->sm_count)
		grid = ctx->sm_count;
	const pbs::Muls mu = { 2u, 4u, 16u };
	k<<<(unsigned) (grid < 1 ? 1 : grid), XW * 32, smem, stream>>>(ctx->d_params, cnt, d_reads, meta, seeds, bins, bins + 2 * pb::PB_SEED_BINS, mu, nullptr, nullptr);
	CUDA_TRY(cudaGetLastError());
	return PB_OK;
}

template <int LML, int LW>
static pb_status overlap_lanes(pb_context *ctx, const LanesState &L, int cnt, int base, const uint32_t *seeds, const int *order, unsigned *bins, cudaStream_t stream) {
	auto k = pbl::assemble_lanes_kernel<LML, LW>;
	constexpr size_t smem = pbl::lanes_smem_bytes<LML, LW>();
	static bool configured[16] = { false };
	if (!configured[ctx->device & 15]) {
		CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
		configured[ctx->device & 15] = true;
	}
	long long grid = (((long long) cnt + 31) / 32 + LW - 1) / LW;
	if (grid > ctx->sm_count)
		grid = ctx->sm_count;
	const size_t nt_row = L.seq_stride / 2;
	k<<<(unsigned) (grid < 1 ? 1 : grid), LW * 32, smem, stream>>>(ctx->d_params, cnt, L.d_reads, L.d_meta + base, seeds, order, L.d_results + base,
	                                                                L.d_seq_nt ? L.d_seq_nt + (size_t) base * nt_row : nullptr, (long long) L.seq_stride, L.d_counters,
	                                                                L.d_list, L.d_count, ctx->d_defer_total, bins + 2 * pb::PB_SEED_BINS + 1, nullptr, base);
	CUDA_TRY(cudaGetLastError());
	return PB_OK;
}

template <int XWO, int LWO>
static pb_status launch_lanes_overlap(pb_context *ctx, int n, const uint8_t *d_reads, const pb_pair_meta *d_meta, pb_pair_result *d_results,
                                      uint8_t *d_seq_nt, size_t seq_stride, unsigned long long *d_counters, cudaStream_t stream, int slices) {
	LanesState L = { ctx, n, 0, d_reads, d_meta, d_results, d_seq_nt, seq_stride, d_counters, stream, nullptr, nullptr, nullptr, 0 };
	PB_TRY(lanes_prepare(L, false));
	PB_TRY(ensure_overlap(ctx));
	cudaStream_t aux = ctx->ovl_stream;
	CUDA_TRY(cudaMemsetAsync(ctx->d_ovl, 0, PB_OVL_SLICES * PB_BINS_STRIDE * sizeof(unsigned), stream));
	const bool timed = ctx->timing && stream == ctx->stream;
	if (timed)
		CUDA_TRY(cudaEventRecord(ctx->tev[0], stream));
	CUDA_TRY(cudaEventRecord(ctx->ovl_ev[PB_OVL_SLICES], stream));
	CUDA_TRY(cudaStreamWaitEvent(aux, ctx->ovl_ev[PB_OVL_SLICES], 0));
	const int per = (((n + slices - 1) / slices) + 31) & ~31;
	constexpr int SW = pb::seed_words(160);
	int i = 0;
	for (int base = 0; base < n; base += per, i++) {
		const int cnt = n - base < per ? n - base : per;
		unsigned *bins = ctx->d_ovl + (size_t) i * PB_BINS_STRIDE;
		uint32_t *seeds = L.d_seeds + (size_t) base * SW;
		int *order = ctx->d_order[0] + base;
		if (i == 0)
			PB_TRY((overlap_sweep<5, 21>(ctx, cnt, d_reads, d_meta + base, seeds, bins, stream)));
		else
			PB_TRY((overlap_sweep<5, XWO>(ctx, cnt, d_reads, d_meta + base, seeds, bins, stream)));
		pb::bin_order_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, stream>>>(cnt, seeds, SW, pb::seed_mask_words(160) + 1, bins, order, nullptr, nullptr);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(ctx->ovl_ev[i], stream));
		CUDA_TRY(cudaStreamWaitEvent(aux, ctx->ovl_ev[i], 0));
		if (base + per >= n)
			PB_TRY((overlap_lanes<152, 12>(ctx, L, cnt, base, seeds, order, bins, aux)));
		else
			PB_TRY((overlap_lanes<152, LWO>(ctx, L, cnt, base, seeds, order, bins, aux)));
	}
	CUDA_TRY(cudaEventRecord(ctx->ovl_ev[PB_OVL_SLICES + 1], aux));
	CUDA_TRY(cudaStreamWaitEvent(stream, ctx->ovl_ev[PB_OVL_SLICES + 1], 0));
	if (timed) {
		CUDA_TRY(cudaEventRecord(ctx->tev[1], stream));
		CUDA_TRY(cudaEventRecord(ctx->tev[2], stream));
	}
	ctx->lanes_pairs += (unsigned long long) n;
	ctx->timing_overlap = true;
	pb_status st = launch_assemble<160, false, 28, false>(ctx, n, d_reads, d_meta, d_results, d_seq_nt, nullptr, seq_stride, d_counters, stream, false, L.d_list, L.d_count);
	ctx->timing_overlap = false;
	return st;
}

static pb_status launch_lanes(pb_context *ctx, int n, int max_len, const uint8_t *d_reads, const pb_pair_meta *d_meta,
                              pb_pair_result *d_results, uint8_t *d_seq_nt, size_t seq_stride,
                              unsigned long long *d_counters, cudaStream_t stream, bool sweep) {
	/* experiment, off unless PANDASEQ_B200_OVERLAP=N (N slices) asks for it: large batches of the 152-nt class on the context's own
	 * stream with the two kernels side by side, slice by slice.  Measured slower than back to back. */
	static int overlap_cfg = -1;
	if (overlap_cfg < 0) {
		const char *env = getenv("PANDASEQ_B200_OVERLAP");
		overlap_cfg = env ? atoi(env) : 0;
		if (overlap_cfg > PB_OVL_SLICES)
			overlap_cfg = PB_OVL_SLICES;
	}
	if (sweep && max_len <= 152 && overlap_cfg >= 2 && n >= (1 << 20) && stream == ctx->stream)
		return launch_lanes_overlap<8, 8>(ctx, n, d_reads, d_meta, d_results, d_seq_nt, seq_stride, d_counters, stream, overlap_cfg);
	LanesState L = { ctx, n, stream == ctx->copy_stream ? 1 : 0, d_reads, d_meta, d_results, d_seq_nt, seq_stride, d_counters, stream, nullptr, nullptr, nullptr, 0 };
	/* one class in batch order while every read fits the 160-nt kernels (or the hash join is asked for, which takes whole
	 * batches of reads up to 256 nt); by length class otherwise */
	const bool classes = sweep && max_len > 160;
	PB_TRY(lanes_prepare(L, classes));
	const bool timed = ctx->timing && stream == ctx->stream;
	if (timed)
		CUDA_TRY(cudaEventRecord(ctx->tev[0], stream));
	/* <seeding class, hash-join warps, sweep warps> / <lane-kernel class, lane warps>: as many warps as the per-warp shared memory
	 * allows; reads up to 152 nt (2x150 included) leave room for a 12th warp of the lane kernel */
	const int top = max_len <= 256 ? 1 : 2;      /* the highest class that can hold a pair of this batch */
	if (classes) {
		pb::class_list_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(n, d_meta, ctx->d_classes[L.si], L.cap, L.class_count(0), L.d_list, L.d_count, ctx->d_defer_total);
		CUDA_TRY(cudaGetLastError());
		PB_TRY((lanes_seed<160, 0, 21>(L, 0, true)));
		PB_TRY((lanes_seed<256, 0, 14>(L, 1, true)));
		if (top >= 2)
			PB_TRY((lanes_seed<320, 0, 10>(L, 2, true)));
	} else if (max_len <= 160) {
		PB_TRY((lanes_seed<160, 32, 21>(L, -1, sweep)));
	} else {
		PB_TRY((lanes_seed<256, 19, 14>(L, -1, sweep)));
	}
	if (timed)
		CUDA_TRY(cudaEventRecord(ctx->tev[1], stream));
	if (classes) {
		PB_TRY((lanes_assemble<160, 11>(L, 0)));
		PB_TRY((lanes_assemble<256, 7>(L, 1)));
		if (top >= 2)
			PB_TRY((lanes_assemble<320, 6>(L, 2)));
	} else if (max_len <= 152) {
		PB_TRY((lanes_assemble<152, 12>(L, -1)));
	} else if (max_len <= 160) {
		PB_TRY((lanes_assemble<160, 11>(L, -1)));
	} else {
		PB_TRY((lanes_assemble<256, 7>(L, -1)));
	}
	if (timed)
		CUDA_TRY(cudaEventRecord(ctx->tev[2], stream));
	ctx->lanes_pairs += (unsigned long long) n;
	/* the pairs handed on, by the general kernel of the batch's length class */
#define PB_LIST(ML, W) return launch_assemble<ML, false, W, false>(ctx, n, d_reads, d_meta, d_results, d_seq_nt, nullptr, seq_stride, d_counters, stream, false, L.d_list, L.d_count)
	if (max_len <= 160) PB_LIST(160, 28);
	if (max_len <= 256) PB_LIST(256, 15);
	if (max_len <= 320) PB_LIST(320, 14);
	PB_LIST(456, 8);
#undef PB_LIST
}

pb_status pb_assemble_dispatch(pb_context *ctx, const pb_config *cfg, int n, int max_len,
                                   const uint8_t *d_reads, const pb_pair_meta *d_meta, pb_pair_result *d_results,
                                   uint8_t *d_seq_nt, double *d_seq_p, size_t seq_stride, unsigned long long *d_counters,
                                   cudaStream_t stream, uint16_t *d_seq_code) {
	const bool over = cfg->algo == PB_RDP_MLE;      /* pear scores from the reconstruction table, only rdp_mle needs its own */
	if (max_len <= 0 || max_len > PB_MAX_LEN)
		max_len = PB_MAX_LEN;
	/* the full kernel carries the primer scans and the log()-based scorers; everything else runs the lean one */
	const bool full0 = cfg->post_primers != 0 || cfg->forward_primer_length > 0 || cfg->reverse_primer_length > 0
		|| cfg->algo == PB_EA_UTIL || cfg->algo == PB_STITCH || cfg->hang_forward_length > 0 || cfg->hang_reverse_length > 0;
	bool stage_seq = cfg->post_primers != 0;
	for (int k = 0; k < cfg->nfilters && k < PB_MAX_FILTERS; k++)
		if (cfg->filters[k].kind == PB_FILTER_MIN_PHRED)
			stage_seq = true;
	const bool full = full0 || stage_seq;
	/* the common case goes lane-per-pair (pb_lanes.cuh); PANDASEQ_B200_LANES=0 keeps everything on the general kernel */
	static int lanes_on = -1;
	if (lanes_on < 0) {
		const char *env = getenv("PANDASEQ_B200_LANES");
		lanes_on = (env && atoi(env) == 0) ? 0 : 1;
	}
	if (d_seq_code && stage_seq) {
		pb_set_error("per-base codes are not available together with primers-after or min_phred (the sequence is staged as doubles)");
		return PB_ERR_ARGUMENT;
	}
	if ((ctx->lanes_mode < 0 ? lanes_on : ctx->lanes_mode) && !full && !d_seq_p && !d_seq_code && cfg->forward_trim == 0 && cfg->reverse_trim == 0
	    && (cfg->algo == PB_SIMPLE_BAYES || cfg->algo == PB_UPARSE || cfg->algo == PB_FLASH || cfg->algo == PB_PEAR) && ((uintptr_t) d_seq_nt % 8) == 0)
	{
		/* seeding: the diagonal sweep (pb_sweep.cuh) unless an explicit maxoverlap lets overlaps run past a read's end, which
		 * only the hash join (pb::seed_kernel) covers; PANDASEQ_B200_SWEEP=0 / pb_set_lanes(ctx, 2) keep the hash join for A/B runs */
		static int sweep_on = -1;
		if (sweep_on < 0) {
			const char *env = getenv("PANDASEQ_B200_SWEEP");
			sweep_on = (env && atoi(env) == 0) ? 0 : 1;
		}
		const bool sweep = sweep_on && ctx->lanes_mode != 2 && cfg->maxoverlap == 0;
		if (sweep || max_len <= 256)
			return launch_lanes(ctx, n, max_len, d_reads, d_meta, d_results, d_seq_nt, seq_stride, d_counters, stream, sweep);
	}
#define PB_GO(ML, OVER, W) do { if (full) return launch_assemble<ML, OVER, W, true>(ctx, n, d_reads, d_meta, d_results, d_seq_nt, d_seq_p, seq_stride, d_counters, stream, stage_seq, nullptr, nullptr, d_seq_code); \
	return launch_assemble<ML, OVER, W, false>(ctx, n, d_reads, d_meta, d_results, d_seq_nt, d_seq_p, seq_stride, d_counters, stream, false, nullptr, nullptr, d_seq_code); } while (0)
	/* warps per CTA: as many as the per-warp shared memory of the class allows next to the LUTs (227 KB per SM) */
	if (max_len <= 160) {
		if (over) PB_GO(160, true, 23);
		else {
			/* 30 warps x 64 registers vs 28 x 72: measured; PANDASEQ_B200_WARPS160 overrides for experiments */
			static int w160 = -1;
			if (w160 < 0) {
				const char *env = getenv("PANDASEQ_B200_WARPS160");
				w160 = env ? atoi(env) : 28;
			}
			if (w160 == 24) PB_GO(160, false, 24);
			else PB_GO(160, false, 28);
		}
	} else if (max_len <= 256) {
		if (over) PB_GO(256, true, 12); else PB_GO(256, false, 15);
	} else if (max_len <= 320) {
		if (over) PB_GO(320, true, 12); else PB_GO(320, false, 14);
	} else {
		if (over) PB_GO(456, true, 6); else PB_GO(456, false, 8);
	}
#undef PB_GO
}

extern "C" pb_status pb_assemble_device(pb_context *ctx, const pb_config *cfg, size_t n, int max_read_len,
                                        const uint8_t *d_reads, const pb_pair_meta *d_meta,
                                        pb_pair_result *d_results, uint8_t *d_seq_nt, double *d_seq_p,
                                        size_t seq_stride, int64_t *d_counters) {
	if (!ctx || !cfg || !d_results || !d_counters || n > 0x7FFFFFFFull || ((d_seq_nt || d_seq_p) && (seq_stride % 16) != 0)) {
		pb_set_error("pb_assemble_device: bad argument (seq_stride must be a multiple of 16)");
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(ctx->device));
	pb_status st = pb_upload_params(ctx, cfg);
	if (st != PB_OK)
		return st;
	if (n == 0)
		return PB_OK;
	return pb_assemble_dispatch(ctx, cfg, (int) n, max_read_len, d_reads, d_meta, d_results, d_seq_nt, d_seq_p, seq_stride,
	                         (unsigned long long *) d_counters, ctx->stream, nullptr);
}

extern "C" pb_status pb_pack_device(pb_context *ctx, size_t n,
                                    const panda_qual *d_f_data, const uint64_t *d_f_off,
                                    const panda_qual *d_r_data, const uint64_t *d_r_off,
                                    const uint32_t *d_rec_off16, uint8_t *d_reads, pb_pair_meta *d_meta) {
	if (!ctx || n > 0x7FFFFFFFull) {
		pb_set_error("pb_pack_device: bad argument");
		return PB_ERR_ARGUMENT;
	}
	if (n == 0)
		return PB_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	const int threads = 256;
	const unsigned blocks = (unsigned) (((long long) n * 32 + threads - 1) / threads);
	pb::pack_kernel<<<blocks, threads, 0, ctx->stream>>>((int) n, (const uint8_t *) d_f_data, (const unsigned long long *) d_f_off, 0ull,
	                                                      (const uint8_t *) d_r_data, (const unsigned long long *) d_r_off, 0ull,
	                                                      d_rec_off16, d_reads, d_meta);
	CUDA_TRY(cudaGetLastError());
	return PB_OK;
}

extern "C" pb_status pb_lanes_stats(pb_context *ctx, uint64_t *lanes_pairs, uint64_t *deferred_pairs) {
	if (!ctx || !lanes_pairs || !deferred_pairs) {
		pb_set_error("pb_lanes_stats: bad argument");
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(ctx->device));
	CUDA_TRY(cudaDeviceSynchronize());
	unsigned long long d = 0;
	CUDA_TRY(cudaMemcpy(&d, ctx->d_defer_total, sizeof d, cudaMemcpyDeviceToHost));
	*lanes_pairs = ctx->lanes_pairs;
	*deferred_pairs = d;
	return PB_OK;
}

extern "C" pb_status pb_set_lanes(pb_context *ctx, int mode) {
	if (!ctx || mode < -1 || mode > 2) {
		pb_set_error("pb_set_lanes: bad argument");
		return PB_ERR_ARGUMENT;
	}
	ctx->lanes_mode = mode;
	return PB_OK;
}

extern "C" pb_status pb_set_timing(pb_context *ctx, int on) {
	if (!ctx) {
		pb_set_error("pb_set_timing: no context");
		return PB_ERR_ARGUMENT;
	}
	ctx->timing = on != 0;
	ctx->timing_kind = 0;
	return PB_OK;
}

extern "C" pb_status pb_last_timing(pb_context *ctx, int *kind, float ms[3]) {
	if (!ctx || !kind || !ms) {
		pb_set_error("pb_last_timing: bad argument");
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(ctx->device));
	ms[0] = ms[1] = ms[2] = 0.0f;
	*kind = ctx->timing_kind;
	if (ctx->timing_kind == 0)
		return PB_OK;
	CUDA_TRY(cudaEventSynchronize(ctx->tev[3]));
	if (ctx->timing_kind == 1) {
		CUDA_TRY(cudaEventElapsedTime(&ms[2], ctx->tev[0], ctx->tev[3]));
	} else {
		CUDA_TRY(cudaEventElapsedTime(&ms[0], ctx->tev[0], ctx->tev[1]));
		CUDA_TRY(cudaEventElapsedTime(&ms[1], ctx->tev[1], ctx->tev[2]));
		CUDA_TRY(cudaEventElapsedTime(&ms[2], ctx->tev[2], ctx->tev[3]));
	}
	return PB_OK;
}

/* ---- host-buffer path ------------------------------------------------------------ */

template <typename T> static cudaError_t regrow_dev(T **p, size_t *cap, size_t need) {
	if (need <= *cap)
		return cudaSuccess;
	cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	cudaError_t e = cudaMalloc((void **) p, need * sizeof(T));
	if (e == cudaSuccess)
		*cap = need;
	return e;
}
template <typename T> static cudaError_t regrow_host(T **p, size_t *cap, size_t need) {
	if (need <= *cap)
		return cudaSuccess;
	cudaFreeHost(*p);
	*p = nullptr;
	*cap = 0;
	cudaError_t e = cudaMallocHost((void **) p, need * sizeof(T));
	if (e == cudaSuccess)
		*cap = need;
	return e;
}

/* Every buffer of a slot has its own capacity and is grown on its own (freed pointer nulled, capacity zeroed first), so a
 * failed allocation leaves the slot consistent. */
static pb_status ensure_slot(pb_context::Slot &s, size_t pairs, size_t fbases, size_t rbases, bool aos_in, bool stage_in,
                             bool stage_res, size_t nt_bytes, size_t p_elems, size_t code_elems) {
	const size_t bases = fbases > rbases ? fbases : rbases;
	const size_t pcap = pairs + pairs / 4 + 16, bcap = bases + bases / 4 + 64;
	CUDA_TRY(regrow_dev(&s.d_meta, &s.cap_meta, pcap));
	CUDA_TRY(regrow_dev(&s.d_res, &s.cap_res, pcap));
	if (aos_in) {
		CUDA_TRY(regrow_host(&s.h_recoff, &s.cap_hrecoff, pcap));
		CUDA_TRY(regrow_dev(&s.d_foff, &s.cap_foff, pcap + 1));
		CUDA_TRY(regrow_dev(&s.d_roff, &s.cap_roff, pcap + 1));
		CUDA_TRY(regrow_dev(&s.d_recoff, &s.cap_recoff, pcap));
		CUDA_TRY(regrow_dev(&s.d_f, &s.cap_f, bcap * 2));
		CUDA_TRY(regrow_dev(&s.d_r, &s.cap_r, bcap * 2));
		if (stage_in) {          /* pinned staging for pageable caller arrays */
			CUDA_TRY(regrow_host(&s.h_foff, &s.cap_hfoff, pcap + 1));
			CUDA_TRY(regrow_host(&s.h_roff, &s.cap_hroff, pcap + 1));
			CUDA_TRY(regrow_host(&s.h_f, &s.cap_hf, bcap * 2));
			CUDA_TRY(regrow_host(&s.h_r, &s.cap_hr, bcap * 2));
		}
	}
	if (stage_res)
		CUDA_TRY(regrow_host(&s.h_res, &s.cap_hres, pcap));
	CUDA_TRY(regrow_host(&s.h_nt, &s.cap_nt, nt_bytes));
	CUDA_TRY(regrow_host(&s.h_p, &s.cap_p, p_elems));
	CUDA_TRY(regrow_host(&s.h_code, &s.cap_code, code_elems));
	return PB_OK;
}

bool pb_is_pinned(const void *p) {
	if (!p)
		return false;
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return attr.type == cudaMemoryTypeHost;
}

static size_t host_chunk_pairs() {
	static pthread_once_t once = PTHREAD_ONCE_INIT;
	static size_t chunk_cfg;
	pthread_once(&once, [] {
		const char *env = getenv("PANDASEQ_B200_CHUNK");      /* pairs per chunk of the host path (experiments) */
		chunk_cfg = env ? (size_t) atol(env) : (size_t) (1u << 18);      /* 256 K pairs: measured best on B200 (82 vs 76 Mpairs/s at 1 M) */
		if (chunk_cfg < 1024)
			chunk_cfg = 1024;
	});
	return chunk_cfg;
}

/* What a host-path call works on.  Input is either the flat AoS arrays of the reference's API (f_data .. r_off; packed on the
 * device) or records the caller packed itself (reads / meta in the layout of include/pandaseq_b200.h: 26 % fewer bytes on
 * the link).  Per-base log p leaves either as doubles (seq_p) or as 16-bit codes into the posterior table (seq_code). */
struct HostJob {
	size_t n;
	const panda_qual *f_data, *r_data;
	const uint64_t *f_off, *r_off;
	const uint8_t *reads;
	const pb_pair_meta *meta;
	int packed_max_len;
	pb_pair_result *results;
	uint8_t *seq_nt;
	double *seq_p;
	uint16_t *seq_code;
	size_t seq_stride;
	int64_t *counters;
};

/* Chunked over two slots, each with its own stream: while chunk k is packed/assembled on the GPU, chunk k+1's
 * host->device copies and chunk k-1's device->host copies run on the other stream.  Caller buffers that are
 * pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are copied from/to directly; pageable
 * buffers go through the slot's pinned staging area first. */
static pb_status assemble_host_chunks(pb_context *ctx, const pb_config *cfg, const HostJob &j) {
	const size_t n = j.n;
	const bool aos = j.reads == nullptr;
	const bool pin_in = aos ? (pb_is_pinned(j.f_data) && pb_is_pinned(j.r_data) && pb_is_pinned(j.f_off) && pb_is_pinned(j.r_off))
	                        : (pb_is_pinned(j.reads) && pb_is_pinned(j.meta));
	const bool pin_res = pb_is_pinned(j.results), pin_nt = pb_is_pinned(j.seq_nt), pin_p = pb_is_pinned(j.seq_p), pin_code = pb_is_pinned(j.seq_code);
	const size_t chunk_cfg = host_chunk_pairs();
	const size_t CHUNK = n > 2 * chunk_cfg ? chunk_cfg : (n + 1) / 2 + 1;       /* at least two chunks so the slots overlap */
	const size_t nt_row = j.seq_stride / 2, stride = j.seq_stride;
	struct Pending { bool live; size_t begin, count; } pend[PB_HOST_SLOTS] = {};
	cudaStream_t streams[2] = { ctx->stream, ctx->copy_stream };
	auto drain = [&](int si) -> pb_status {
		if (!pend[si].live)
			return PB_OK;
		pb_context::Slot &s = ctx->slot[si];
		pend[si].live = false;
		CUDA_TRY(cudaEventSynchronize(s.done));
		if (!pin_res)
			memcpy(j.results + pend[si].begin, s.h_res, pend[si].count * sizeof(pb_pair_result));
		if (j.seq_nt && !pin_nt)
			memcpy(j.seq_nt + pend[si].begin * nt_row, s.h_nt, pend[si].count * nt_row);
		if (j.seq_p && !pin_p)
			memcpy(j.seq_p + pend[si].begin * stride, s.h_p, pend[si].count * stride * sizeof(double));
		if (j.seq_code && !pin_code)
			memcpy(j.seq_code + pend[si].begin * stride, s.h_code, pend[si].count * stride * sizeof(uint16_t));
		return PB_OK;
	};
	/* packed input: every read length is checked before anything is queued */
	if (!aos) {
		for (size_t i = 0; i < n; i++)
			if (j.meta[i].flen != 0xFFFFu && (j.meta[i].flen > PB_MAX_LEN || j.meta[i].rlen > PB_MAX_LEN)) {
				pb_set_error("read longer than PANDA_MAX_LEN (pair %zu)", i);
				return PB_ERR_ARGUMENT;
			}
	}
	int si = 0;
	for (size_t begin = 0; begin < n; begin += CHUNK, si = (si + 1) % PB_HOST_SLOTS) {
		const size_t count = (n - begin < CHUNK) ? (n - begin) : CHUNK;
		pb_status st = drain(si);
		if (st != PB_OK)
			return st;
		pb_context::Slot &s = ctx->slot[si];
		cudaStream_t stream = streams[si & 1];
		size_t max_len = 0, total16 = 0, fbases = 0, rbases = 0;
		uint64_t fb = 0, rb = 0;
		if (aos) {
			fb = j.f_off[begin];
			rb = j.r_off[begin];
			fbases = (size_t) (j.f_off[begin + count] - fb);
			rbases = (size_t) (j.r_off[begin + count] - rb);
		}
		st = ensure_slot(s, count, fbases, rbases, aos, !pin_in, !pin_res, (j.seq_nt && !pin_nt) ? count * nt_row : 0,
		                 (j.seq_p && !pin_p) ? count * stride : 0, (j.seq_code && !pin_code) ? count * stride : 0);
		if (st != PB_OK)
			return st;
		if (aos) {
			/* layout: record offsets (integer bookkeeping only) and the longest read of the chunk */
			for (size_t i = 0; i < count; i++) {
				const size_t fl = (size_t) (j.f_off[begin + i + 1] - j.f_off[begin + i]), rl = (size_t) (j.r_off[begin + i + 1] - j.r_off[begin + i]);
				if (fl > max_len) max_len = fl;
				if (rl > max_len) max_len = rl;
				s.h_recoff[i] = (uint32_t) total16;
				total16 += pb_record_bytes(fl, rl) / 16;
			}
			if (max_len > PB_MAX_LEN) {
				pb_set_error("read longer than PANDA_MAX_LEN (%zu > %d)", max_len, PB_MAX_LEN);
				return PB_ERR_ARGUMENT;
			}
		} else {
			/* the chunk's records are contiguous in the caller's buffer: [off16 of its first pair, end of its last pair) */
			const pb_pair_meta &first = j.meta[begin], &last = j.meta[begin + count - 1];
			const size_t last_bytes = last.flen == 0xFFFFu ? 0 : pb_record_bytes(last.flen, last.rlen);
			total16 = (size_t) last.off16 + last_bytes / 16 - first.off16;
			max_len = (size_t) j.packed_max_len;
		}
		CUDA_TRY(regrow_dev(&s.d_reads, &s.cap_reads, total16 * 16 + 16));
		if (j.seq_nt)
			CUDA_TRY(regrow_dev(&s.d_nt, &s.cap_dnt, count * nt_row));
		if (j.seq_p)
			CUDA_TRY(regrow_dev(&s.d_p, &s.cap_dp, count * stride));
		if (j.seq_code)
			CUDA_TRY(regrow_dev(&s.d_code, &s.cap_dcode, count * stride));
		if (aos) {
			const void *src_f = j.f_data + fb, *src_r = j.r_data + rb, *src_fo = j.f_off + begin, *src_ro = j.r_off + begin;
			if (!pin_in) {
				memcpy(s.h_f, src_f, fbases * 2);
				memcpy(s.h_r, src_r, rbases * 2);
				memcpy(s.h_foff, src_fo, (count + 1) * 8);
				memcpy(s.h_roff, src_ro, (count + 1) * 8);
				src_f = s.h_f; src_r = s.h_r; src_fo = s.h_foff; src_ro = s.h_roff;
			}
			CUDA_TRY(cudaMemcpyAsync(s.d_f, src_f, fbases * 2, cudaMemcpyHostToDevice, stream));
			CUDA_TRY(cudaMemcpyAsync(s.d_r, src_r, rbases * 2, cudaMemcpyHostToDevice, stream));
			CUDA_TRY(cudaMemcpyAsync(s.d_foff, src_fo, (count + 1) * 8, cudaMemcpyHostToDevice, stream));
			CUDA_TRY(cudaMemcpyAsync(s.d_roff, src_ro, (count + 1) * 8, cudaMemcpyHostToDevice, stream));
			CUDA_TRY(cudaMemcpyAsync(s.d_recoff, s.h_recoff, count * 4, cudaMemcpyHostToDevice, stream));
			const int threads = 256;
			const unsigned blocks = (unsigned) (((long long) count * 32 + threads - 1) / threads);
			pb::pack_kernel<<<blocks, threads, 0, stream>>>((int) count, s.d_f, s.d_foff, fb, s.d_r, s.d_roff, rb, s.d_recoff, s.d_reads, s.d_meta);
			CUDA_TRY(cudaGetLastError());
		} else {
			const uint8_t *src_reads = j.reads + (size_t) j.meta[begin].off16 * 16;
			const pb_pair_meta *src_meta = j.meta + begin;
			if (!pin_in) {
				CUDA_TRY(regrow_host(&s.h_reads, &s.cap_hreads, total16 * 16));
				CUDA_TRY(regrow_host(&s.h_meta, &s.cap_hmeta, count));
				memcpy(s.h_reads, src_reads, total16 * 16);
				memcpy(s.h_meta, src_meta, count * sizeof(pb_pair_meta));
				src_reads = s.h_reads;
				src_meta = s.h_meta;
			}
			CUDA_TRY(cudaMemcpyAsync(s.d_reads, src_reads, total16 * 16, cudaMemcpyHostToDevice, stream));
			CUDA_TRY(cudaMemcpyAsync(s.d_meta, src_meta, count * sizeof(pb_pair_meta), cudaMemcpyHostToDevice, stream));
			if (j.meta[begin].off16 != 0) {      /* offsets are relative to the caller's buffer: rebase them on the chunk */
				const unsigned blocks = (unsigned) ((count + 255) / 256);
				pb::rebase_meta_kernel<<<blocks, 256, 0, stream>>>((int) count, s.d_meta, j.meta[begin].off16);
				CUDA_TRY(cudaGetLastError());
			}
		}
		st = pb_assemble_dispatch(ctx, cfg, (int) count, (int) max_len, s.d_reads, s.d_meta, s.d_res,
		                       j.seq_nt ? s.d_nt : nullptr, j.seq_p ? s.d_p : nullptr, stride, ctx->d_counters, stream,
		                       j.seq_code ? s.d_code : nullptr);
		if (st != PB_OK)
			return st;
		CUDA_TRY(cudaMemcpyAsync(pin_res ? (void *) (j.results + begin) : (void *) s.h_res, s.d_res, count * sizeof(pb_pair_result), cudaMemcpyDeviceToHost, stream));
		if (j.seq_nt)
			CUDA_TRY(cudaMemcpyAsync(pin_nt ? (void *) (j.seq_nt + begin * nt_row) : (void *) s.h_nt, s.d_nt, count * nt_row, cudaMemcpyDeviceToHost, stream));
		if (j.seq_p)
			CUDA_TRY(cudaMemcpyAsync(pin_p ? (void *) (j.seq_p + begin * stride) : (void *) s.h_p, s.d_p, count * stride * sizeof(double), cudaMemcpyDeviceToHost, stream));
		if (j.seq_code)
			CUDA_TRY(cudaMemcpyAsync(pin_code ? (void *) (j.seq_code + begin * stride) : (void *) s.h_code, s.d_code, count * stride * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream));
		CUDA_TRY(cudaEventRecord(s.done, stream));
		pend[si].live = true;
		pend[si].begin = begin;
		pend[si].count = count;
	}
	for (int k = 0; k < PB_HOST_SLOTS; k++) {
		pb_status st = drain(k);
		if (st != PB_OK)
			return st;
	}
	return PB_OK;
}

static pb_status assemble_host_locked(pb_context *ctx, const pb_config *cfg, const HostJob &j) {
	const bool aos = j.reads == nullptr;
	if (!cfg || (!j.results && j.n) || (j.n && aos && (!j.f_data || !j.f_off || !j.r_data || !j.r_off)) || (j.n && !aos && !j.meta)
	    || ((j.seq_nt || j.seq_p || j.seq_code) && (j.seq_stride % 16) != 0) || (j.seq_p && j.seq_code)) {
		pb_set_error("pb_assemble_host: bad argument (seq_stride must be a multiple of 16; per-base log p as doubles or as codes, not both)");
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(ctx->device));
	pb_status st = pb_upload_params(ctx, cfg);
	if (st != PB_OK)
		return st;
	if (j.n == 0)
		return PB_OK;
	CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, PB_NCOUNTERS * sizeof(unsigned long long), ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));      /* parameters + zeroed counters visible to both streams */
	st = assemble_host_chunks(ctx, cfg, j);
	if (st != PB_OK) {
		/* chunks queued before the failure may still be copying from / into the caller's buffers and the slots' staging:
		 * nothing is left in flight when the call returns */
		cudaStreamSynchronize(ctx->stream);
		cudaStreamSynchronize(ctx->copy_stream);
		return st;
	}
	if (j.counters) {
		unsigned long long hc[PB_NCOUNTERS];
		CUDA_TRY(cudaMemcpy(hc, ctx->d_counters, sizeof hc, cudaMemcpyDeviceToHost));
		int64_t tmp[PB_NCOUNTERS];
		for (int i = 0; i < PB_NCOUNTERS; i++)
			tmp[i] = (int64_t) hc[i];
		pb_counters_merge(j.counters, tmp);
	}
	return PB_OK;
}

static pb_status assemble_host_entry(pb_context *ctx, const pb_config *cfg, const HostJob &j, const char *who) {
	if (!ctx) {
		pb_set_error("%s: no context", who);
		return PB_ERR_ARGUMENT;
	}
	pthread_mutex_lock(&ctx->lock);
	pb_status st = assemble_host_locked(ctx, cfg, j);
	pthread_mutex_unlock(&ctx->lock);
	return st;
}

extern "C" pb_status pb_assemble_host(pb_context *ctx, const pb_config *cfg, size_t n,
                                      const panda_qual *f_data, const uint64_t *f_off,
                                      const panda_qual *r_data, const uint64_t *r_off,
                                      pb_pair_result *results, uint8_t *seq_nt, double *seq_p,
                                      size_t seq_stride, int64_t *counters) {
	const HostJob j = { n, f_data, r_data, f_off, r_off, nullptr, nullptr, 0, results, seq_nt, seq_p, nullptr, seq_stride, counters };
	return assemble_host_entry(ctx, cfg, j, "pb_assemble_host");
}

extern "C" pb_status pb_assemble_host_codes(pb_context *ctx, const pb_config *cfg, size_t n,
                                            const panda_qual *f_data, const uint64_t *f_off,
                                            const panda_qual *r_data, const uint64_t *r_off,
                                            pb_pair_result *results, uint8_t *seq_nt, uint16_t *seq_code,
                                            size_t seq_stride, int64_t *counters) {
	const HostJob j = { n, f_data, r_data, f_off, r_off, nullptr, nullptr, 0, results, seq_nt, nullptr, seq_code, seq_stride, counters };
	return assemble_host_entry(ctx, cfg, j, "pb_assemble_host_codes");
}

extern "C" pb_status pb_assemble_host_packed(pb_context *ctx, const pb_config *cfg, size_t n, int max_read_len,
                                             const uint8_t *reads, const pb_pair_meta *meta,
                                             pb_pair_result *results, uint8_t *seq_nt, size_t seq_stride, int64_t *counters) {
	if (n && !reads) {
		pb_set_error("pb_assemble_host_packed: no records");
		return PB_ERR_ARGUMENT;
	}
	const HostJob j = { n, nullptr, nullptr, nullptr, nullptr, reads, meta, max_read_len <= 0 ? PB_MAX_LEN : max_read_len,
	                    results, seq_nt, nullptr, nullptr, seq_stride, counters };
	return assemble_host_entry(ctx, cfg, j, "pb_assemble_host_packed");
}

/* Page-locked host memory for callers that have no CUDA runtime of their own (the panda_* object layer is plain C). */
extern "C" void *pb_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}
extern "C" void pb_host_free(void *p) {
	if (p)
		cudaFreeHost(p);
}

/* process-wide contexts for the panda_* object layer: one per GPU, made on first use.  The default one is the device
 * PANDASEQ_B200_DEVICE names (0 if unset). */
static pb_context *g_device_ctx[64];
static pthread_mutex_t g_shared_lock = PTHREAD_MUTEX_INITIALIZER;

extern "C" pb_status pb_device_context(int device, pb_context **out) {
	pb_status st = PB_OK;
	if (device < 0 || device >= 64) {
		pb_set_error("device %d out of range", device);
		return PB_ERR_ARGUMENT;
	}
	pthread_mutex_lock(&g_shared_lock);
	if (!g_device_ctx[device])
		st = pb_context_create(device, &g_device_ctx[device]);
	*out = g_device_ctx[device];
	pthread_mutex_unlock(&g_shared_lock);
	return st;
}

/* The calling thread onto the host cores of the socket a GPU hangs off (sysfs: local_cpulist of its PCI device), so that the
 * page-locked staging it allocates and fills afterwards lies in that socket's memory and the copies do not cross the socket link.
 * Best effort: without the sysfs entry, or where the process's cpuset leaves no such core, nothing changes. */
extern "C" void pb_bind_thread_near_device(int device) {
	char bus[32], path[128], list[4096];
	if (cudaDeviceGetPCIBusId(bus, (int) sizeof bus, device) != cudaSuccess) {
		cudaGetLastError();
		return;
	}
	for (char *c = bus; *c; c++)
		if (*c >= 'A' && *c <= 'F')
			*c = (char) (*c - 'A' + 'a');
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
	FILE *f = fopen(path, "r");
	if (!f)
		return;
	const bool got = fgets(list, (int) sizeof list, f) != nullptr;
	fclose(f);
	if (!got)
		return;
	cpu_set_t allowed, want;
	if (sched_getaffinity(0, sizeof allowed, &allowed) != 0)
		return;
	CPU_ZERO(&want);
	int picked = 0;
	for (char *p = list; *p && *p != '\n';) {
		char *end;
		long a = strtol(p, &end, 10), b = a;
		if (end == p)
			break;
		if (*end == '-')
			b = strtol(end + 1, &end, 10);
		for (long c = a; c <= b && c < CPU_SETSIZE; c++)
			if (CPU_ISSET((int) c, &allowed)) {
				CPU_SET((int) c, &want);
				picked++;
			}
		p = *end == ',' ? end + 1 : end;
	}
	if (picked > 0)
		pthread_setaffinity_np(pthread_self(), sizeof want, &want);
}

extern "C" pb_status pb_shared_context(pb_context **out) {
	int dev = 0;
	const char *env = getenv("PANDASEQ_B200_DEVICE");
	if (env)
		dev = atoi(env);
	return pb_device_context(dev, out);
}
