/* pb_ctx.h -- the device context shared by pb_device.cu (assembly) and pb_io.cu (FASTQ in / text out). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pb_internal.h"

struct pb_io_state;

/* chunks of the host path in flight, alternating between the two streams (4 were measured: no gain, PCIe is the limit) */
#define PB_HOST_SLOTS 2

struct pb_context {
	int device;
	int sm_count;
	cudaStream_t stream;
	cudaStream_t copy_stream;
	pb_device_params *d_params;      /* in HBM */
	pb_device_params *h_params;      /* pinned mirror */
	pb_config cached_cfg;
	bool cfg_valid;
	unsigned long long *d_counters;  /* scratch counters for the host path */
	uint8_t *d_scratch;              /* per-warp scratch of the primers-after path, allocated on first use */
	size_t scratch_bytes;
	/* deferral lists of the lane-per-pair kernel, one per stream of the host path: [0] = count, [1 .. 3] = pairs per length class, [8 ..] = pair indices */
	int *d_defer[2];
	size_t defer_cap[2];
	uint32_t *d_seeds[2];            /* candidate-overlap masks of pb::seed_kernel, 8 words per pair */
	int *d_order[2];                 /* the pairs of a launch bin by bin (pb::bin_order_kernel) */
	unsigned *d_bins[2];             /* per length class: 2 x PB_SEED_BINS (pairs per bin, cursors) + the kernels' batch counters */
	/* the two kernels side by side on slices of a large batch (pb_device.cu, launch_lanes_overlap): second stream, events, per-slice bin state */
	cudaStream_t ovl_stream;
	cudaEvent_t ovl_ev[20];
	unsigned *d_ovl;
	bool ovl_ready, timing_overlap;
	int *d_classes[2];               /* the pairs of a batch listed by length class (pb::class_list_kernel); allocated on first use */
	bool classes_on[2];
	double *d_pear_cdf;              /* pear_test table (PB_PEAR_ROWS x PB_PEAR_COLS), built on first use */
	unsigned long long *d_defer_total;   /* pairs deferred so far (device), next to lanes_pairs (host) */
	unsigned long long lanes_pairs;
	int lanes_mode;                  /* -1 = follow PANDASEQ_B200_LANES (default on), 0 / 1 = pb_set_lanes */
	/* pb_set_timing / pb_last_timing: events around the kernels of the last pb_assemble_device call */
	bool timing;
	int timing_kind;                 /* 0 = nothing recorded, 1 = general kernel alone, 2 = seed + lanes + general(list) */
	cudaEvent_t tev[4];
	pthread_mutex_t lock;            /* one host-path call at a time per context (assemblers share the process-wide one) */
	/* host-path staging (grown on demand) */
	struct Slot {
		/* every buffer with its own capacity (elements) */
		uint8_t *h_f, *h_r;              /* pinned AoS staging */
		size_t cap_hf, cap_hr;
		unsigned long long *h_foff, *h_roff;
		size_t cap_hfoff, cap_hroff;
		uint32_t *h_recoff;
		size_t cap_hrecoff;
		uint8_t *d_f, *d_r;
		size_t cap_f, cap_r;
		unsigned long long *d_foff, *d_roff;
		size_t cap_foff, cap_roff;
		uint32_t *d_recoff;
		size_t cap_recoff;
		uint8_t *d_reads;
		size_t cap_reads;
		pb_pair_meta *d_meta;
		size_t cap_meta;
		pb_pair_result *d_res, *h_res;
		size_t cap_res, cap_hres;
		uint8_t *d_nt, *h_nt;
		double *d_p, *h_p;
		uint16_t *d_code, *h_code;       /* per-base posterior codes (pb_assemble_host_codes) */
		pb_pair_meta *h_meta;            /* pinned staging of caller-packed input (pb_assemble_host_packed) */
		uint8_t *h_reads;
		size_t cap_nt, cap_p, cap_dnt, cap_dp, cap_code, cap_dcode, cap_hmeta, cap_hreads;
		cudaEvent_t done;
	} slot[PB_HOST_SLOTS];
	pb_io_state *io;                 /* buffers of the FASTQ / text stages, allocated on first use (pb_io.cu) */
};

#define CUDA_TRY(expr)                                                                         \
	do {                                                                                       \
		cudaError_t e_ = (expr);                                                               \
		if (e_ != cudaSuccess) {                                                               \
			pb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
			return PB_ERR_CUDA;                                                                \
		}                                                                                      \
	} while (0)


pb_status pb_assemble_dispatch(pb_context *ctx, const pb_config *cfg, int n, int max_len,
                               const uint8_t *d_reads, const pb_pair_meta *d_meta, pb_pair_result *d_results,
                               uint8_t *d_seq_nt, double *d_seq_p, size_t seq_stride, unsigned long long *d_counters,
                               cudaStream_t stream, uint16_t *d_seq_code = nullptr);
pb_status pb_upload_params(pb_context *ctx, const pb_config *cfg);
bool pb_is_pinned(const void *p);
void pb_io_release(pb_context *ctx);
