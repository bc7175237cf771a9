/* pb_io.cu -- host side of the FASTQ-in / text-out stages: launch wrappers for pb_io.cuh and the
 * host-buffer chain FASTQ text -> parse -> assemble -> FASTA/FASTQ text (pb_fastq_assemble_host).
 *
 * As in pb_device.cu, the host only moves bytes and does integer bookkeeping; every byte of read data
 * is interpreted on the GPU.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pb_ctx.h"
#include "pb_io.cuh"

using pbio::ParseState;
using pbio::TextView;

struct IoSlot {
	/* text in */
	uint8_t *d_text[2];
	size_t cap_text[2];
	uint8_t *h_text[2];            /* pinned staging for pageable caller buffers */
	size_t cap_htext[2];
	/* line index */
	uint32_t *d_nl[2];
	size_t cap_nl[2];
	uint32_t *d_blk[2];
	size_t cap_blk[2];
	ParseState *d_state, *h_state;
	/* parsed records */
	uint8_t *d_reads;
	size_t cap_reads;
	pb_pair_meta *d_meta;
	pb_seq_id *d_ids;
	size_t cap_rec;
	/* results */
	pb_pair_result *d_res;
	uint8_t *d_nt;
	double *d_p;
	size_t cap_res, cap_nt, cap_p;
	/* text out */
	uint32_t *d_len;
	unsigned long long *d_off, *d_tile, *d_total, *h_total;
	size_t cap_len, cap_tile;
	char *d_out;
	size_t cap_out;
	char *h_out;
	size_t cap_hout;
	cudaEvent_t ev_parse, ev_fmt, ev_done;
};

struct pb_io_state {
	IoSlot slot[2];
};

template <typename T> static cudaError_t grow_dev(T **p, size_t *cap, size_t need, size_t slack_div = 4) {
	if (need <= *cap)
		return cudaSuccess;
	cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	const size_t want = need + need / slack_div + 64;
	cudaError_t e = cudaMalloc((void **) p, want * sizeof(T));
	if (e == cudaSuccess)
		*cap = want;
	return e;
}
template <typename T> static cudaError_t grow_host(T **p, size_t *cap, size_t need) {
	if (need <= *cap)
		return cudaSuccess;
	cudaFreeHost(*p);
	*p = nullptr;
	*cap = 0;
	const size_t want = need + need / 4 + 64;
	cudaError_t e = cudaMallocHost((void **) p, want * sizeof(T));
	if (e == cudaSuccess)
		*cap = want;
	return e;
}

static pb_status io_get(pb_context *ctx, pb_io_state **out) {
	if (!ctx->io) {
		pb_io_state *io = (pb_io_state *) calloc(1, sizeof(pb_io_state));
		if (!io)
			return PB_ERR_NOMEM;
		for (int k = 0; k < 2; k++) {
			IoSlot &s = io->slot[k];
			CUDA_TRY(cudaMalloc(&s.d_state, sizeof(ParseState)));
			CUDA_TRY(cudaMallocHost(&s.h_state, sizeof(ParseState)));
			CUDA_TRY(cudaMalloc(&s.d_total, sizeof(unsigned long long)));
			CUDA_TRY(cudaMallocHost(&s.h_total, sizeof(unsigned long long)));
			CUDA_TRY(cudaEventCreateWithFlags(&s.ev_parse, cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&s.ev_fmt, cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
		}
		ctx->io = io;
	}
	*out = ctx->io;
	return PB_OK;
}

void pb_io_release(pb_context *ctx) {
	if (!ctx->io)
		return;
	for (int k = 0; k < 2; k++) {
		IoSlot &s = ctx->io->slot[k];
		for (int f = 0; f < 2; f++) {
			cudaFree(s.d_text[f]);
			cudaFreeHost(s.h_text[f]);
			cudaFree(s.d_nl[f]);
			cudaFree(s.d_blk[f]);
		}
		cudaFree(s.d_state); cudaFreeHost(s.h_state);
		cudaFree(s.d_reads); cudaFree(s.d_meta); cudaFree(s.d_ids);
		cudaFree(s.d_res); cudaFree(s.d_nt); cudaFree(s.d_p);
		cudaFree(s.d_len); cudaFree(s.d_off); cudaFree(s.d_tile); cudaFree(s.d_total); cudaFreeHost(s.h_total);
		cudaFree(s.d_out); cudaFreeHost(s.h_out);
		cudaEventDestroy(s.ev_parse); cudaEventDestroy(s.ev_fmt); cudaEventDestroy(s.ev_done);
	}
	free(ctx->io);
	ctx->io = nullptr;
}

/* ---- parse ---------------------------------------------------------------------------------------------- */
static pb_status parse_launch(pb_context *ctx, IoSlot &s, cudaStream_t st, const uint8_t *d_fwd, size_t fb, const uint8_t *d_rev, size_t rb,
                              int qualmin, int policy, uint8_t *d_reads, size_t reads_cap, pb_pair_meta *d_meta, pb_seq_id *d_ids,
                              size_t max_records) {
	const size_t bytes[2] = { fb, rb };
	const uint8_t *text[2] = { d_fwd, d_rev };
	TextView tv[2];
	for (int f = 0; f < 2; f++) {
		const size_t nblocks = (bytes[f] + pbio::NL_TILE - 1) / pbio::NL_TILE;
		/* room for one newline per 16 bytes, and never fewer than the records asked for; an index that does not fit
		 * is reported (nl_overflow) and the caller retries with a full-size one */
		size_t want_nl = bytes[f] / 16 + 1024;
		if (want_nl < 4 * max_records + 8)
			want_nl = 4 * max_records + 8;
		if (want_nl > bytes[f] + 8)
			want_nl = bytes[f] + 8;
		if (s.cap_nl[f] < want_nl)
			CUDA_TRY(grow_dev(&s.d_nl[f], &s.cap_nl[f], want_nl));
		CUDA_TRY(grow_dev(&s.d_blk[f], &s.cap_blk[f], nblocks + 1));
		tv[f].text = text[f];
		tv[f].bytes = bytes[f];
		tv[f].nl = s.d_nl[f];
		tv[f].nl_cap = (unsigned) (s.cap_nl[f] > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : s.cap_nl[f]);
		tv[f].block_cnt = s.d_blk[f];
		tv[f].nblocks = (unsigned) nblocks;
	}
	ParseState init;
	memset(&init, 0, sizeof init);
	init.err_key = pbio::NO_ERROR_KEY;
	*s.h_state = init;
	CUDA_TRY(cudaMemcpyAsync(s.d_state, s.h_state, sizeof(ParseState), cudaMemcpyHostToDevice, st));
	const unsigned nbmax = tv[0].nblocks > tv[1].nblocks ? tv[0].nblocks : tv[1].nblocks;
	if (nbmax > 0) {
		pbio::nl_count<<<dim3(nbmax, 2), pbio::NL_THREADS, 0, st>>>(tv[0], tv[1]);
		pbio::nl_scan<<<2, 1024, 0, st>>>(tv[0], tv[1], s.d_state);
		pbio::nl_write<<<dim3(nbmax, 2), pbio::NL_THREADS, 0, st>>>(tv[0], tv[1]);
		const int grid = ctx->sm_count * 8;
		pbio::fq_geometry<<<grid, 256, 0, st>>>(tv[0], tv[1], s.d_state, (unsigned) max_records);
		pbio::fq_stride<<<1, 1, 0, st>>>(s.d_state);
		pbio::fq_ids<<<ctx->sm_count * 6, pbio::ID_THREADS, 0, st>>>(tv[0], tv[1], s.d_state, policy, d_ids);
		pbio::fq_reads<5><<<ctx->sm_count * 16, 128, 0, st>>>(tv[0], tv[1], s.d_state, qualmin, d_reads, (unsigned long long) reads_cap, d_meta);
		pbio::fq_reads<10><<<ctx->sm_count * 10, 128, 0, st>>>(tv[0], tv[1], s.d_state, qualmin, d_reads, (unsigned long long) reads_cap, d_meta);
		pbio::fq_reads<15><<<ctx->sm_count * 6, 128, 0, st>>>(tv[0], tv[1], s.d_state, qualmin, d_reads, (unsigned long long) reads_cap, d_meta);
		pbio::fq_finish<<<1, 1024, 0, st>>>(s.d_state, d_meta);
		CUDA_TRY(cudaGetLastError());
	}
	CUDA_TRY(cudaMemcpyAsync(s.h_state, s.d_state, sizeof(ParseState), cudaMemcpyDeviceToHost, st));
	return PB_OK;
}

static void fill_info(const ParseState &ps, pb_fastq_info *info) {
	memset(info, 0, sizeof *info);
	info->records = ps.records;
	info->limit = ps.limit;
	info->pairs = ps.pairs;
	info->consumed_fwd = ps.consumed[0];
	info->consumed_rev = ps.consumed[1];
	info->error = ps.error;
	info->max_read_len = (int32_t) (ps.max_len[0] > ps.max_len[1] ? ps.max_len[0] : ps.max_len[1]);
	info->stride16 = ps.stride16;
}

extern "C" pb_status pb_fastq_parse_device(pb_context *ctx, const char *d_fwd, size_t fwd_bytes, const char *d_rev, size_t rev_bytes,
                                           int qualmin, int policy, size_t max_records,
                                           uint8_t *d_reads, size_t reads_capacity, pb_pair_meta *d_meta, pb_seq_id *d_ids, pb_fastq_info *info) {
	if (!ctx || !info || !d_meta || !d_reads || fwd_bytes > 0xFFFFFFF0ull || rev_bytes > 0xFFFFFFF0ull || max_records > 0x7FFFFFFFull
	    || ((uintptr_t) d_fwd & 15) || ((uintptr_t) d_rev & 15)) {
		pb_set_error("pb_fastq_parse_device: bad argument (texts must be 16-byte aligned and below 4 GiB each)");
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(ctx->device));
	pthread_mutex_lock(&ctx->lock);
	pb_io_state *io;
	pb_status rc = io_get(ctx, &io);
	if (rc == PB_OK) {
		IoSlot &s = io->slot[0];
		for (int attempt = 0; attempt < 2 && rc == PB_OK; attempt++) {
			rc = parse_launch(ctx, s, ctx->stream, (const uint8_t *) d_fwd, fwd_bytes, (const uint8_t *) d_rev, rev_bytes, qualmin, policy,
			                  d_reads, reads_capacity, d_meta, d_ids, max_records);
			if (rc != PB_OK)
				break;
			if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
				pb_set_error("FASTQ parse failed: %s", cudaGetErrorString(cudaGetLastError()));
				rc = PB_ERR_CUDA;
				break;
			}
			if (!s.h_state->nl_overflow)
				break;
			/* pathological text (a newline every few bytes): index at full size and go again */
			for (int f = 0; f < 2 && rc == PB_OK; f++) {
				const size_t full = (f ? rev_bytes : fwd_bytes) + 8;
				if (grow_dev(&s.d_nl[f], &s.cap_nl[f], full, 1u << 30) != cudaSuccess)
					rc = PB_ERR_NOMEM;
			}
		}
		if (rc == PB_OK) {
			fill_info(*s.h_state, info);
			if ((size_t) info->records * info->stride16 * 16 > reads_capacity) {
				pb_set_error("pb_fastq_parse_device: %llu records of %u bytes do not fit d_reads (%zu bytes)",
				             (unsigned long long) info->records, info->stride16 * 16, reads_capacity);
				rc = PB_ERR_ARGUMENT;
			}
		}
	}
	pthread_mutex_unlock(&ctx->lock);
	return rc;
}

/* ---- format --------------------------------------------------------------------------------------------- */
static pb_status ensure_offsets(IoSlot &s, size_t n) {
	/* d_off is sized with d_len (cap_len counts records) */
	static_assert(sizeof(unsigned long long) == 8, "");
	if (!s.d_off || s.cap_len < n + 1) {
		cudaFree(s.d_off);
		s.d_off = nullptr;
		CUDA_TRY(grow_dev(&s.d_len, &s.cap_len, n + 1));
		CUDA_TRY(cudaMalloc(&s.d_off, s.cap_len * sizeof(unsigned long long)));
	}
	return PB_OK;
}

static pb_status format_launch(pb_context *ctx, IoSlot &s, cudaStream_t st, int format, size_t n, const pb_pair_result *d_res,
                               const uint8_t *d_nt, const double *d_p, size_t seq_stride, const pb_seq_id *d_ids, const uint8_t *d_fwd,
                               char *d_text, size_t capacity) {
	const size_t ntiles = (n + pbio::SCAN_TILE - 1) / pbio::SCAN_TILE;      /* d_len / d_off: ensure_offsets() */
	CUDA_TRY(grow_dev(&s.d_tile, &s.cap_tile, ntiles + 1));
	if (n == 0) {
		CUDA_TRY(cudaMemsetAsync(s.d_total, 0, sizeof(unsigned long long), st));
	} else {
		pbio::fmt_length<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>((int) n, format == PB_OUT_FASTQ, d_res, d_ids, s.d_len);
		pbio::scan_tile_sums<<<(unsigned) ntiles, 256, 0, st>>>((int) n, s.d_len, s.d_tile);
		pbio::scan_tiles<<<1, 1024, 0, st>>>((int) ntiles, s.d_tile, s.d_total);
		pbio::scan_apply<<<(unsigned) ntiles, 256, 0, st>>>((int) n, s.d_len, s.d_tile, s.d_off);
		if (d_text && capacity)
			pbio::fmt_write<<<(unsigned) ((n * 32 + 255) / 256), 256, 0, st>>>((int) n, format == PB_OUT_FASTQ, d_res, d_nt, d_p, (long long) seq_stride,
			                                                                        d_ids, d_fwd, s.d_len, s.d_off, ctx->d_params->score, d_text,
			                                                                        (unsigned long long) capacity);
		CUDA_TRY(cudaGetLastError());
	}
	CUDA_TRY(cudaMemcpyAsync(s.h_total, s.d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	return PB_OK;
}

extern "C" pb_status pb_format_device(pb_context *ctx, int format, size_t n, const pb_pair_result *d_results, const uint8_t *d_seq_nt,
                                      const double *d_seq_p, size_t seq_stride, const pb_seq_id *d_ids, const char *d_fwd,
                                      char *d_text, size_t capacity, size_t *text_bytes) {
	if (!ctx || !text_bytes || n > 0x7FFFFFFFull || (n && (!d_results || !d_seq_nt || !d_ids || !d_fwd)) || (format != PB_OUT_FASTA && format != PB_OUT_FASTQ)
	    || (format == PB_OUT_FASTQ && n && !d_seq_p) || (seq_stride % 16) != 0) {
		pb_set_error("pb_format_device: bad argument (FASTQ output needs the per-base log p)");
		return PB_ERR_ARGUMENT;
	}
	if (!ctx->cfg_valid) {
		pb_set_error("pb_format_device: no configuration uploaded yet (assemble first)");
		return PB_ERR_ARGUMENT;
	}
	CUDA_TRY(cudaSetDevice(ctx->device));
	pthread_mutex_lock(&ctx->lock);
	pb_io_state *io;
	pb_status rc = io_get(ctx, &io);
	if (rc == PB_OK)
		rc = ensure_offsets(io->slot[0], n);
	if (rc == PB_OK)
		rc = format_launch(ctx, io->slot[0], ctx->stream, format, n, d_results, d_seq_nt, d_seq_p, seq_stride, d_ids, (const uint8_t *) d_fwd,
		                   d_text, capacity);
	if (rc == PB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
		pb_set_error("format failed: %s", cudaGetErrorString(cudaGetLastError()));
		rc = PB_ERR_CUDA;
	}
	if (rc == PB_OK)
		*text_bytes = (size_t) *io->slot[0].h_total;
	pthread_mutex_unlock(&ctx->lock);
	return rc;
}

/* ---- identifiers on the host (bookkeeping: copies substrings, prints integers) ------------------------------- */
extern "C" void pb_seq_id_expand(const pb_seq_id *id, const char *fwd_text, panda_seq_identifier *out) {
	const char *h = fwd_text + id->hdr_off;
	memset(out, 0, sizeof *out);
	if (id->fmt == PB_IDFMT_SRA || id->fmt == PB_IDFMT_EBI_SRA) {
		snprintf(out->instrument, sizeof out->instrument, "%cRR%d", id->fmt == PB_IDFMT_SRA ? 'S' : 'E', id->sra);
	} else {
		memcpy(out->instrument, h + id->inst_off, id->inst_len < 99 ? id->inst_len : 99);
		if (id->inst_len >= 100)      /* a 100-character field fills the member without a terminator, as in the reference */
			memcpy(out->instrument, h + id->inst_off, 100);
	}
	memcpy(out->run, h + id->run_off, id->run_len > 100 ? 100 : id->run_len);
	memcpy(out->flowcell, h + id->fc_off, id->fc_len > 100 ? 100 : id->fc_len);
	memcpy(out->tag, h + id->tag_off, id->tag_len > PANDA_TAG_LEN ? PANDA_TAG_LEN : id->tag_len);
	out->lane = id->lane;
	out->tile = id->tile;
	out->x = id->x;
	out->y = id->y;
}

/* ---- the host-buffer chain ---------------------------------------------------------------------------------- */
static pb_status stage_text(IoSlot &s, int f, cudaStream_t st, const char *src, size_t bytes, bool pinned) {
	CUDA_TRY(grow_dev(&s.d_text[f], &s.cap_text[f], bytes + 64));
	const void *from = src;
	if (!pinned) {
		CUDA_TRY(grow_host(&s.h_text[f], &s.cap_htext[f], bytes + 64));
		memcpy(s.h_text[f], src, bytes);
		from = s.h_text[f];
	}
	if (bytes)
		CUDA_TRY(cudaMemcpyAsync(s.d_text[f], from, bytes, cudaMemcpyHostToDevice, st));
	return PB_OK;
}

/* everything that is one entry per FASTQ record */
static pb_status ensure_records(IoSlot &s, size_t recs) {
	if (recs <= s.cap_rec && s.d_meta && s.d_ids && s.d_res)
		return PB_OK;
	cudaFree(s.d_meta); cudaFree(s.d_ids); cudaFree(s.d_res);
	s.d_meta = nullptr; s.d_ids = nullptr; s.d_res = nullptr;
	s.cap_rec = 0;
	const size_t cap = recs + recs / 4 + 64;
	CUDA_TRY(cudaMalloc(&s.d_meta, cap * sizeof(pb_pair_meta)));
	CUDA_TRY(cudaMalloc(&s.d_ids, cap * sizeof(pb_seq_id)));
	CUDA_TRY(cudaMalloc(&s.d_res, cap * sizeof(pb_pair_result)));
	s.cap_rec = cap;
	return PB_OK;
}

/* Chunks of both texts go through two slots, each with its own stream.  Per chunk k (slot k & 1):
 *   H2D text -> line index -> ids + reads -> [host learns records / limit / bytes consumed] -> assemble -> format
 *   -> [host learns the text length] -> D2H text.
 * While the host waits for chunk k's parse, chunk k-1 assembles on the other stream; chunk k-1's text comes back while
 * chunk k assembles.  The next window starts where the parse says the last complete record ended. */
static pb_status fastq_assemble_locked(pb_context *ctx, const pb_config *cfg, int qualmin, int policy, int out_format,
                                       const char *fwd, size_t fwd_bytes, const char *rev, size_t rev_bytes, int final,
                                       char *out_text, size_t out_capacity, int64_t *counters, pb_stream_info *info) {
	memset(info, 0, sizeof *info);
	CUDA_TRY(cudaSetDevice(ctx->device));
	pb_status st = pb_upload_params(ctx, cfg);
	if (st != PB_OK)
		return st;
	pb_io_state *io;
	st = io_get(ctx, &io);
	if (st != PB_OK)
		return st;
	CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, PB_NCOUNTERS * sizeof(unsigned long long), ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	cudaStream_t streams[2] = { ctx->stream, ctx->copy_stream };
	const bool pin_f = pb_is_pinned(fwd), pin_r = pb_is_pinned(rev), pin_out = pb_is_pinned(out_text);
	const bool want_p = out_format == PB_OUT_FASTQ;

	size_t win = 96u << 20;                    /* bytes of each text per chunk */
	{
		const char *env = getenv("PANDASEQ_B200_FASTQ_CHUNK");
		if (env && atol(env) > 0)
			win = (size_t) atol(env);
	}
	if (win > 0xF0000000ull)
		win = 0xF0000000ull;
	size_t win_f = win, win_r = win;
	size_t pf = 0, pr = 0, out_pos = 0;
	struct Pending { bool live, finished; size_t out_at; } pend[2] = { { false, false, 0 }, { false, false, 0 } };
	int error = PB_FQ_OK;
	bool stop = false;

	auto finish = [&](int si) -> pb_status {          /* wait for a chunk's text length, start bringing the text back */
		if (!pend[si].live || pend[si].finished)
			return PB_OK;
		IoSlot &s = io->slot[si];
		CUDA_TRY(cudaEventSynchronize(s.ev_fmt));
		const size_t bytes = (size_t) *s.h_total;
		if (out_text && bytes) {
			if (out_pos + bytes > out_capacity) {
				pb_set_error("pb_fastq_assemble_host: output needs more than %zu bytes", out_capacity);
				return PB_ERR_ARGUMENT;
			}
			if (pin_out) {
				CUDA_TRY(cudaMemcpyAsync(out_text + out_pos, s.d_out, bytes, cudaMemcpyDeviceToHost, streams[si]));
			} else {
				CUDA_TRY(grow_host(&s.h_out, &s.cap_hout, bytes));
				CUDA_TRY(cudaMemcpyAsync(s.h_out, s.d_out, bytes, cudaMemcpyDeviceToHost, streams[si]));
			}
		}
		CUDA_TRY(cudaEventRecord(s.ev_done, streams[si]));
		pend[si].out_at = out_pos;
		pend[si].finished = true;
		out_pos += bytes;
		return PB_OK;
	};
	auto drain = [&](int si) -> pb_status {           /* after this the slot's buffers are free again */
		if (!pend[si].live)
			return PB_OK;
		pb_status rc = finish(si);
		if (rc != PB_OK)
			return rc;
		IoSlot &s = io->slot[si];
		CUDA_TRY(cudaEventSynchronize(s.ev_done));
		if (out_text && !pin_out)
			memcpy(out_text + pend[si].out_at, s.h_out, (size_t) *s.h_total);
		pend[si].live = pend[si].finished = false;
		return PB_OK;
	};

	int si = 0, prev = -1;
	while (!stop && pf < fwd_bytes && pr < rev_bytes) {
		st = drain(si);
		if (st != PB_OK)
			return st;
		IoSlot &s = io->slot[si];
		cudaStream_t stream = streams[si];
		const size_t rem_f = fwd_bytes - pf, rem_r = rev_bytes - pr;
		const size_t fb = rem_f < win_f ? rem_f : win_f, rb = rem_r < win_r ? rem_r : win_r;
		const bool at_end = fb == rem_f && rb == rem_r;
		st = stage_text(s, 0, stream, fwd + pf, fb, pin_f);
		if (st == PB_OK)
			st = stage_text(s, 1, stream, rev + pr, rb, pin_r);
		/* room for the records of this window: an estimate (a record of 2 x 100 nt is ~250 bytes of text); the parse
		 * reports what it really found and a denser text gets a second pass with exact sizes */
		if (st == PB_OK)
			st = ensure_records(s, (fb < rb ? fb : rb) / 192 + 1024);
		if (st != PB_OK)
			return st;
		CUDA_TRY(grow_dev(&s.d_reads, &s.cap_reads, fb + rb + (1u << 20)));
		ParseState ps;
		for (int attempt = 0;; attempt++) {
			st = parse_launch(ctx, s, stream, s.d_text[0], fb, s.d_text[1], rb, qualmin, policy, s.d_reads, s.cap_reads, s.d_meta, s.d_ids, s.cap_rec);
			if (st != PB_OK)
				return st;
			CUDA_TRY(cudaEventRecord(s.ev_parse, stream));
			if (prev >= 0) {          /* while this chunk is copied and indexed: the previous chunk's text length and D2H */
				st = finish(prev);
				if (st != PB_OK)
					return st;
			}
			CUDA_TRY(cudaEventSynchronize(s.ev_parse));
			ps = *s.h_state;
			const size_t nl_min = ps.nl_total[0] < ps.nl_total[1] ? ps.nl_total[0] : ps.nl_total[1];
			const size_t recs = nl_min / 4;
			const size_t need_reads = recs * (size_t) ps.stride16 * 16;
			if (!ps.nl_overflow && recs <= s.cap_rec && need_reads <= s.cap_reads)
				break;
			if (attempt >= 2) {
				pb_set_error("FASTQ parse: buffers still too small after resizing");
				return PB_ERR_NOMEM;
			}
			CUDA_TRY(cudaStreamSynchronize(stream));
			if (ps.nl_overflow)
				for (int f = 0; f < 2; f++)
					CUDA_TRY(grow_dev(&s.d_nl[f], &s.cap_nl[f], (f ? rb : fb) + 8, 1u << 30));
			st = ensure_records(s, recs + 1);
			if (st != PB_OK)
				return st;
			CUDA_TRY(grow_dev(&s.d_reads, &s.cap_reads, (recs + 1) * pb_record_bytes(PB_MAX_LEN, PB_MAX_LEN), 1u << 30));
		}
		if (ps.records == 0) {
			if (!at_end) {            /* not one complete record in the window: widen it */
				win_f *= 2;
				win_r *= 2;
				continue;
			}
			if (final && ps.nl_total[0] >= 1 && ps.nl_total[1] >= 1)
				error = PB_FQ_PREMATURE_EOF;      /* a truncated last record (fastq.c:58,69,84) */
			break;
		}
		const size_t limit = (size_t) ps.limit;
		if (ps.error != PB_FQ_OK) {
			error = ps.error;
			stop = true;
		}
		/* assemble + format records [0, limit) of this chunk */
		const int max_len = (int) (ps.max_len[0] > ps.max_len[1] ? ps.max_len[0] : ps.max_len[1]);
		size_t seq_stride = ((size_t) ps.max_len[0] + ps.max_len[1] + 15) & ~(size_t) 15;
		if (seq_stride == 0)
			seq_stride = 16;
		CUDA_TRY(grow_dev(&s.d_nt, &s.cap_nt, (limit + 1) * (seq_stride / 2)));
		if (want_p)
			CUDA_TRY(grow_dev(&s.d_p, &s.cap_p, (limit + 1) * seq_stride));
		st = ensure_offsets(s, limit);
		if (st != PB_OK)
			return st;
		/* an emitted record is never longer than its two input records plus punctuation */
		CUDA_TRY(grow_dev(&s.d_out, &s.cap_out, (size_t) ps.consumed[0] + (size_t) ps.consumed[1] + 48 * limit + 1024));
		if (limit > 0) {
			st = pb_assemble_dispatch(ctx, cfg, (int) limit, max_len, s.d_reads, s.d_meta, s.d_res, s.d_nt, want_p ? s.d_p : nullptr, seq_stride,
			                          ctx->d_counters, stream);
			if (st != PB_OK)
				return st;
		}
		st = format_launch(ctx, s, stream, out_format, limit, s.d_res, s.d_nt, want_p ? s.d_p : nullptr, seq_stride, s.d_ids, s.d_text[0],
		                   out_text ? s.d_out : nullptr, s.cap_out);
		if (st != PB_OK)
			return st;
		CUDA_TRY(cudaEventRecord(s.ev_fmt, stream));
		pend[si].live = true;
		pend[si].finished = false;
		info->records += limit;
		info->pairs += ps.pairs;
		pf += (size_t) ps.consumed[0];
		pr += (size_t) ps.consumed[1];
		if (at_end) {
			if (final && error == PB_FQ_OK && ps.nl_total[0] - 4 * ps.records >= 1 && ps.nl_total[1] - 4 * ps.records >= 1)
				error = PB_FQ_PREMATURE_EOF;
			stop = true;
		} else {
			/* keep the two windows at the same number of records */
			const double per_f = (double) ps.consumed[0] / (double) ps.records, per_r = (double) ps.consumed[1] / (double) ps.records;
			const double target = (double) win / (per_f > per_r ? per_f : per_r);
			win_f = (size_t) (target * per_f) + 8192;
			win_r = (size_t) (target * per_r) + 8192;
		}
		prev = si;
		si ^= 1;
	}
	if (prev >= 0) {          /* text comes back in chunk order */
		st = drain(prev ^ 1);
		if (st != PB_OK)
			return st;
		st = drain(prev);
		if (st != PB_OK)
			return st;
	}
	info->consumed_fwd = pf;
	info->consumed_rev = pr;
	info->out_bytes = out_pos;
	info->error = error;
	if (counters) {
		unsigned long long hc[PB_NCOUNTERS];
		CUDA_TRY(cudaMemcpy(hc, ctx->d_counters, sizeof hc, cudaMemcpyDeviceToHost));
		int64_t tmp[PB_NCOUNTERS];
		for (int i = 0; i < PB_NCOUNTERS; i++)
			tmp[i] = (int64_t) hc[i];
		pb_counters_merge(counters, tmp);
	}
	return PB_OK;
}

extern "C" pb_status pb_fastq_assemble_host(pb_context *ctx, const pb_config *cfg, int qualmin, int policy, int out_format,
                                            const char *fwd, size_t fwd_bytes, const char *rev, size_t rev_bytes, int final,
                                            char *out_text, size_t out_capacity, int64_t *counters, pb_stream_info *info) {
	if (!ctx || !cfg || !info || (fwd_bytes && !fwd) || (rev_bytes && !rev) || (out_format != PB_OUT_FASTA && out_format != PB_OUT_FASTQ)) {
		pb_set_error("pb_fastq_assemble_host: bad argument");
		return PB_ERR_ARGUMENT;
	}
	pthread_mutex_lock(&ctx->lock);
	pb_status st = fastq_assemble_locked(ctx, cfg, qualmin, policy, out_format, fwd, fwd_bytes, rev, rev_bytes, final, out_text, out_capacity, counters, info);
	if (st != PB_OK) {
		cudaStreamSynchronize(ctx->stream);
		cudaStreamSynchronize(ctx->copy_stream);
	}
	pthread_mutex_unlock(&ctx->lock);
	return st;
}
