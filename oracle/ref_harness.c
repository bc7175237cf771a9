/* ref_harness.c -- drives the UNMODIFIED reference (oracle/_ref/libpandaseq_ref.so)
 * over a flat batch, through the reference's own public API
 * (panda_assembler_assemble, assembler.c:368-383), and writes the same
 * structure-of-arrays output as oracle/panda_oracle.c.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by `make ref` against the headers in
 * /root/reference; the built oracle/_ref/libref_harness.so travels to the GPU box,
 * the reference sources do not.  One PandaAssembler per thread, created the way
 * diff.c:180 does (no reader, null-writer logger), panda_debug_flags = 0 so the
 * per-pair "INFO BESTOLP" formatting stays out of the timing.
 */
#include <math.h>
#include <stdio.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pandaseq.h>
#include "panda_oracle.h"

typedef struct {
	const po_config *cfg;
	size_t begin, end;
	const po_qual *f_data;
	const uint64_t *f_off;
	const po_qual *r_data;
	const uint64_t *r_off;
	po_flat_out *out;
	int64_t counters[PO_NCOUNTERS];
	int failed;
} ref_job;

bool no_n_check(PandaLogProxy logger, const panda_result_seq *sequence, void *user_data);      /* args_assembler.c:106 */
bool short_check(PandaLogProxy logger, const panda_result_seq *sequence, void *user_data);     /* args_assembler.c:233 */
bool long_check(PandaLogProxy logger, const panda_result_seq *sequence, void *user_data);      /* args_assembler.c:268 */
#define OPENER(name) bool name##_LTX_opener(PandaLogProxy logger, const char *args, PandaPreCheck *precheck, PandaCheck *check, void **user_data, PandaDestroy *destroy)
OPENER(min_overlapbits);
OPENER(completely_miss_the_point);
OPENER(min_phred);
OPENER(pear_test);

typedef struct {
	long rejected[8];
	int n;
} reject_counts;

static bool count_rejected(PandaAssembler assembler, PandaModule module, size_t rejected, void *data) {
	reject_counts *rc = data;
	(void) assembler;
	(void) module;
	if (rc->n < 8)
		rc->rejected[rc->n++] = (long) rejected;
	return true;
}

/* A log proxy over a writer that discards everything.  panda_log_proxy_new takes its own reference on the writer
 * (proxy.c:44), so the creator's reference is dropped right away, as the reference's own callers do (proxy.c:50-52):
 * a leaked PandaWriter keeps a pthread_key (writer.c:93), and a process has only 1024 of them. */
static PandaLogProxy quiet_logger(void) {
	PandaWriter w = panda_writer_new_null();
	PandaLogProxy logger = panda_log_proxy_new(w);
	panda_writer_unref(w);
	return logger;
}

static PandaAssembler make_assembler(const po_config *cfg) {
	PandaLogProxy logger = quiet_logger();
	PandaAssembler a = panda_assembler_new_kmer(NULL, NULL, NULL, logger, (size_t) cfg->num_kmers);
	PandaAlgorithm algo = NULL;
	panda_log_proxy_unref(logger);
	if (a == NULL)
		return NULL;
	switch (cfg->algo) {
	case PO_SIMPLE_BAYES:
		algo = panda_algorithm_simple_bayes_new();
		panda_algorithm_simple_bayes_set_error_estimation(algo, cfg->sb_q);
		break;
	case PO_PEAR:
		algo = panda_algorithm_pear_new();
		panda_algorithm_pear_set_random_base_log_p(algo, cfg->pear_random_base);
		break;
	case PO_RDP_MLE:
		algo = panda_algorithm_rdp_mle_new();
		break;
	case PO_FLASH:
		algo = panda_algorithm_flash_new();
		break;
	case PO_EA_UTIL:
		algo = panda_algorithm_ea_util_new();
		break;
	case PO_STITCH:
		algo = panda_algorithm_stitch_new();
		break;
	case PO_UPARSE:
		algo = panda_algorithm_uparse_new();
		panda_algorithm_uparse_set_error_estimation(algo, cfg->sb_q);
		break;
	}
	if (algo == NULL) {
		panda_assembler_unref(a);
		return NULL;
	}
	panda_assembler_set_algorithm(a, algo);
	panda_algorithm_unref(algo);
	{
		/* The setter takes a probability and stores log(p) (assembler_support.c:384-390);
		 * find the p whose log is bit-identical to the requested log-threshold. */
		double want = cfg->threshold, p = exp(want), lo = p, hi = p;
		int found = (log(p) == want);
		for (int k = 0; k < 8 && !found; k++) {
			lo = nextafter(lo, 0.0);
			hi = nextafter(hi, 2.0);
			if (log(lo) == want) { p = lo; found = 1; }
			else if (log(hi) == want) { p = hi; found = 1; }
		}
		if (!found) {
			panda_assembler_unref(a);
			return NULL;
		}
		panda_assembler_set_threshold(a, p);
	}
	panda_assembler_set_minimum_overlap(a, (int) cfg->minoverlap);
	panda_assembler_set_maximum_overlap(a, (int) cfg->maxoverlap);
	panda_assembler_set_primer_penalty(a, cfg->primer_penalty);
	panda_assembler_set_primers_after(a, cfg->post_primers != 0);
	if (cfg->forward_primer_length > 0)
		panda_assembler_set_forward_primer(a, (panda_nt *) cfg->forward_primer, (size_t) cfg->forward_primer_length);
	else
		panda_assembler_set_forward_trim(a, (size_t) cfg->forward_trim);
	if (cfg->reverse_primer_length > 0)
		panda_assembler_set_reverse_primer(a, (panda_nt *) cfg->reverse_primer, (size_t) cfg->reverse_primer_length);
	else
		panda_assembler_set_reverse_trim(a, (size_t) cfg->reverse_trim);
	/* module_checkseq (module.c:124-137) with the reference's own check functions: the three built into the library
	 * (args_assembler.c) and four plugins compiled into this harness from their sources (oracle/Makefile) */
	PandaLogProxy quiet = quiet_logger();
	for (int k = 0; k < cfg->nfilters && k < 7; k++) {
		const struct po_filter *f = &cfg->filters[k];
		PandaModule m = NULL;
		char args[128];
		PandaPreCheck precheck = NULL;
		PandaCheck check = NULL;
		void *user = NULL;
		PandaDestroy destroy = NULL;
		switch (f->kind) {
		case PO_FILTER_NO_N: m = panda_module_new("N", no_n_check, NULL, NULL, NULL); break;
		case PO_FILTER_SHORT: m = panda_module_new("SHORT", short_check, NULL, (void *) (size_t) f->ivalue, NULL); break;
		case PO_FILTER_LONG: m = panda_module_new("LONG", long_check, NULL, (void *) (size_t) f->ivalue, NULL); break;
		case PO_FILTER_MIN_OVERLAPBITS:
			snprintf(args, sizeof args, "%.17g", f->dvalue);
			if (min_overlapbits_LTX_opener(quiet, args, &precheck, &check, &user, &destroy))
				m = panda_module_new("min_overlapbits", check, precheck, user, destroy);
			break;
		case PO_FILTER_MISS_THE_POINT:
			snprintf(args, sizeof args, "%d", f->ivalue);
			if (completely_miss_the_point_LTX_opener(quiet, args, &precheck, &check, &user, &destroy))
				m = panda_module_new("completely_miss_the_point", check, precheck, user, destroy);
			break;
		case PO_FILTER_MIN_PHRED:
			snprintf(args, sizeof args, "%d", f->ivalue);
			if (min_phred_LTX_opener(quiet, args, &precheck, &check, &user, &destroy))
				m = panda_module_new("min_phred", check, precheck, user, destroy);
			break;
		case PO_FILTER_PEAR_TEST:
			/* The plugin cannot be given other values than its defaults: panda_parse_key_values (args.c:613-633) rejects a
			 * string with a second key, and the opener hands the LOGGER, not its parameter struct, to the key processor
			 * (plugin_pear_test.c:96), so a single key overwrites the logger and leaves the parameter untouched.  The
			 * reference therefore always tests with alpha = 1, beta = -1, cutoff = 0.01; only that is pinned here. */
			if (f->dvalue == 1.0 && f->dvalue2 == -1.0 && f->dvalue3 == 0.01
			    && pear_test_LTX_opener(quiet, NULL, &precheck, &check, &user, &destroy))
				m = panda_module_new("pear_test", check, precheck, user, destroy);
			break;
		}
		if (m == NULL) {
			panda_log_proxy_unref(quiet);
			panda_assembler_unref(a);
			return NULL;
		}
		panda_assembler_add_module(a, m);
		panda_module_unref(m);
	}
	panda_log_proxy_unref(quiet);
	return a;
}

/* The reference's threshold setter takes a probability and stores its log
 * (assembler_support.c:384-390); exp(log x) may not round-trip bit-exactly, so the
 * harness also exposes what the reference ended up with. */
double ref_effective_threshold(const po_config *cfg) {
	PandaAssembler a = make_assembler(cfg);
	double t;
	if (a == NULL)
		return 0;
	t = log(panda_assembler_get_threshold(a));
	panda_assembler_unref(a);
	return t;
}

static void blank_row(po_flat_out *o, size_t i, int status) {
	if (o->status) o->status[i] = (uint8_t) status;
	if (o->slow) o->slow[i] = 0;
	if (o->overlap) o->overlap[i] = 0;
	if (o->seq_len) o->seq_len[i] = 0;
	if (o->mismatches) o->mismatches[i] = 0;
	if (o->degenerates) o->degenerates[i] = 0;
	if (o->examined) o->examined[i] = 0;
	if (o->fwd_offset) o->fwd_offset[i] = 0;
	if (o->rev_offset) o->rev_offset[i] = 0;
	if (o->quality) o->quality[i] = 0;
	if (o->est_prob) o->est_prob[i] = 0;
	if (o->seq_nt) memset(o->seq_nt + i * (size_t) o->seq_stride, 0, (size_t) o->seq_stride);
	if (o->seq_p) memset(o->seq_p + i * (size_t) o->seq_stride, 0, (size_t) o->seq_stride * sizeof(double));
}

static void process_one(ref_job *job, PandaAssembler a, size_t i, const panda_qual *F, size_t flen, const panda_qual *R, size_t rlen) {
	po_flat_out *o = job->out;
	panda_seq_identifier id;
	panda_qual fpad[PO_MAX_LEN];
	reject_counts before, after;
	memset(&id, 0, sizeof id);
	memset(&before, 0, sizeof before);
	memset(&after, 0, sizeof after);
	{
		long slow_before = panda_assembler_get_slow_count(a);
		long lowq_before = panda_assembler_get_low_quality_count(a);
		long badr_before = panda_assembler_get_bad_read_count(a);
		long nofp_before = panda_assembler_get_no_forward_primer_count(a);
		long norp_before = panda_assembler_get_no_reverse_primer_count(a);
		long noalgn_before = panda_assembler_get_failed_alignment_count(a);
		panda_assembler_foreach_module(a, count_rejected, &before);
		if (job->cfg->algo == PO_PEAR) {
			/* algo_pear.c:52,54 read forward[rindex]; pin the out-of-range case to zeros. */
			memset(fpad, 0, sizeof fpad);
			memcpy(fpad, F, flen * sizeof(panda_qual));
			F = fpad;
		}
		const panda_result_seq *res = panda_assembler_assemble(a, &id, F, flen, R, rlen);
		int status;
		if (res != NULL)
			status = PO_OK;
		else if (panda_assembler_get_low_quality_count(a) != lowq_before)
			status = PO_LOWQ;
		else if (panda_assembler_get_bad_read_count(a) != badr_before)
			status = PO_BADR;
		else if (panda_assembler_get_no_forward_primer_count(a) != nofp_before)
			status = PO_NOFP;
		else if (panda_assembler_get_no_reverse_primer_count(a) != norp_before)
			status = PO_NORP;
		else if (panda_assembler_get_failed_alignment_count(a) != noalgn_before)
			status = PO_NOALGN;
		else {
			status = PO_NOALGN;
			panda_assembler_foreach_module(a, count_rejected, &after);
			for (int k = 0; k < after.n; k++)
				if (after.rejected[k] != before.rejected[k])
					status = PO_FILTERED + k;
		}
		if (o->status) o->status[i] = (uint8_t) status;
		if (o->slow) o->slow[i] = (uint8_t) (panda_assembler_get_slow_count(a) != slow_before);
		/* On LOWQ the reference returns NULL but has filled its result; that object is
		 * not reachable through the public API, so only OK rows carry the fields. */
		int emitted = (res != NULL);
		if (o->overlap) o->overlap[i] = emitted ? (int32_t) res->overlap : 0;
		if (o->seq_len) o->seq_len[i] = emitted ? (int32_t) res->sequence_length : 0;
		if (o->mismatches) o->mismatches[i] = emitted ? (int32_t) res->overlap_mismatches : 0;
		if (o->degenerates) o->degenerates[i] = emitted ? (int32_t) res->degenerates : 0;
		if (o->examined) o->examined[i] = emitted ? (int32_t) res->overlaps_examined : 0;
		if (o->fwd_offset) o->fwd_offset[i] = emitted ? (int32_t) res->forward_offset : 0;
		if (o->rev_offset) o->rev_offset[i] = emitted ? (int32_t) res->reverse_offset : 0;
		if (o->quality) o->quality[i] = emitted ? res->quality : 0;
		if (o->est_prob) o->est_prob[i] = emitted ? res->estimated_overlap_probability : 0;
		if (o->seq_nt) {
			uint8_t *dst = o->seq_nt + i * (size_t) o->seq_stride;
			memset(dst, 0, (size_t) o->seq_stride);
			if (emitted && (size_t) o->seq_stride >= res->sequence_length)
				for (size_t k = 0; k < res->sequence_length; k++)
					dst[k] = (uint8_t) res->sequence[k].nt;
		}
		if (o->seq_p) {
			double *dst = o->seq_p + i * (size_t) o->seq_stride;
			memset(dst, 0, (size_t) o->seq_stride * sizeof(double));
			if (emitted && (size_t) o->seq_stride >= res->sequence_length)
				for (size_t k = 0; k < res->sequence_length; k++)
					dst[k] = res->sequence[k].p;
		}
	}
}

/* the pairs of a job as a PandaNextSeq, so that the reference's own panda_trim_overhangs can sit on top (hang.c:82-113) */
typedef struct {
	ref_job *job;
	size_t cur;
} pair_source;

static bool pair_source_next(panda_seq_identifier *id, const panda_qual **forward, size_t *forward_length,
                             const panda_qual **reverse, size_t *reverse_length, void *user) {
	pair_source *src = user;
	ref_job *job = src->job;
	if (src->cur >= job->end)
		return false;
	size_t i = src->cur++;
	memset(id, 0, sizeof *id);
	*forward = (const panda_qual *) (job->f_data + job->f_off[i]);
	*forward_length = job->f_off[i + 1] - job->f_off[i];
	*reverse = (const panda_qual *) (job->r_data + job->r_off[i]);
	*reverse_length = job->r_off[i + 1] - job->r_off[i];
	return true;
}

static void *run_job(void *arg) {
	ref_job *job = arg;
	PandaAssembler a = make_assembler(job->cfg);
	memset(job->counters, 0, sizeof job->counters);
	if (a == NULL) {
		job->failed = 1;
		return NULL;
	}
	if (job->cfg->hang_forward_length > 0 || job->cfg->hang_reverse_length > 0) {
		pair_source src = { job, job->begin };
		void *next_data = NULL;
		PandaDestroy next_destroy = NULL;
		PandaLogProxy logger = quiet_logger();
		PandaNextSeq next = panda_trim_overhangs(pair_source_next, &src, NULL, logger, (panda_nt *) job->cfg->hang_forward,
		                                         (size_t) job->cfg->hang_forward_length, (panda_nt *) job->cfg->hang_reverse,
		                                         (size_t) job->cfg->hang_reverse_length, job->cfg->hang_skip != 0, job->cfg->hang_threshold,
		                                         &next_data, &next_destroy);
		for (;;) {
			panda_seq_identifier id;
			const panda_qual *F, *R;
			size_t flen, rlen, before = src.cur;
			bool more = next(&id, &F, &flen, &R, &rlen, next_data);
			size_t upto = more ? src.cur - 1 : job->end;
			for (size_t i = before; i < upto; i++)
				blank_row(job->out, i, PO_SKIP);       /* dropped by the trimmer: the assembler never sees them */
			if (!more)
				break;
			process_one(job, a, src.cur - 1, F, flen, R, rlen);
		}
		next_destroy(next_data);
		panda_log_proxy_unref(logger);
	} else {
		for (size_t i = job->begin; i < job->end; i++)
			process_one(job, a, i, (const panda_qual *) (job->f_data + job->f_off[i]), job->f_off[i + 1] - job->f_off[i],
			            (const panda_qual *) (job->r_data + job->r_off[i]), job->r_off[i + 1] - job->r_off[i]);
	}
	job->counters[PO_C_COUNT] = panda_assembler_get_count(a);
	job->counters[PO_C_OK] = panda_assembler_get_ok_count(a);
	job->counters[PO_C_LOWQ] = panda_assembler_get_low_quality_count(a);
	job->counters[PO_C_NOALGN] = panda_assembler_get_failed_alignment_count(a);
	job->counters[PO_C_BADR] = panda_assembler_get_bad_read_count(a);
	job->counters[PO_C_NOFP] = panda_assembler_get_no_forward_primer_count(a);
	job->counters[PO_C_NORP] = panda_assembler_get_no_reverse_primer_count(a);
	job->counters[PO_C_SLOW] = panda_assembler_get_slow_count(a);
	job->counters[PO_C_LONGEST] = (int64_t) panda_assembler_get_longest_overlap(a);
	{
		reject_counts rc;
		memset(&rc, 0, sizeof rc);
		panda_assembler_foreach_module(a, count_rejected, &rc);
		for (int k = 0; k < rc.n && k < 7; k++)
			job->counters[PO_C_REJECTED + k] = rc.rejected[k];
	}
	for (size_t k = 0; k < 2 * PO_MAX_LEN; k++)
		job->counters[PO_C_OVERLAPS + k] = panda_assembler_get_overlap_count(a, k);
	panda_assembler_unref(a);
	return NULL;
}

int ref_assemble_flat(const po_config *cfg, size_t n,
                      const po_qual *f_data, const uint64_t *f_off,
                      const po_qual *r_data, const uint64_t *r_off,
                      po_flat_out *out, int threads) {
	panda_debug_flags = 0;
	if (threads < 1)
		threads = 1;
	if ((size_t) threads > n)
		threads = n ? (int) n : 1;
	ref_job *jobs = calloc((size_t) threads, sizeof(ref_job));
	pthread_t *tids = calloc((size_t) threads, sizeof(pthread_t));
	for (int k = 0; k < threads; k++) {
		jobs[k].cfg = cfg;
		jobs[k].begin = n * (size_t) k / (size_t) threads;
		jobs[k].end = n * (size_t) (k + 1) / (size_t) threads;
		jobs[k].f_data = f_data;
		jobs[k].f_off = f_off;
		jobs[k].r_data = r_data;
		jobs[k].r_off = r_off;
		jobs[k].out = out;
		if (k > 0)
			pthread_create(&tids[k], NULL, run_job, &jobs[k]);
	}
	run_job(&jobs[0]);
	int failed = jobs[0].failed;
	for (int k = 1; k < threads; k++) {
		pthread_join(tids[k], NULL);
		failed |= jobs[k].failed;
	}
	if (out->counters) {
		for (int k = 0; k < threads; k++)
			for (int c = 0; c < PO_NCOUNTERS; c++) {
				if (c == PO_C_LONGEST) {
					if (out->counters[c] < jobs[k].counters[c])
						out->counters[c] = jobs[k].counters[c];
				} else {
					out->counters[c] += jobs[k].counters[c];
				}
			}
	}
	free(jobs);
	free(tids);
	return failed ? -1 : 0;
}

/* Raw access to the reference's generated LUTs (table.c) for the table parity test. */
extern const double qual_match_simple_bayesian[][47];
extern const double qual_mismatch_simple_bayesian[][47];
extern const double qual_match_pear[][47];
extern const double qual_mismatch_pear[][47];
extern const double qual_mismatch_rdp_mle[][47];
extern const double qual_mismatch_assembled_rdp_mle[][47];
extern const double qual_match_uparse[][47];
extern const double qual_mismatch_uparse[][47];
extern const double qual_score[47];
extern const double qual_score_err[47];

void ref_get_tables(po_tables *t) {
	t->qual_nn = panda_algorithm_simple_bayes_class.prob_unpaired;
	memcpy(t->match_sb, qual_match_simple_bayesian, sizeof t->match_sb);
	memcpy(t->mismatch_sb, qual_mismatch_simple_bayesian, sizeof t->mismatch_sb);
	memcpy(t->match_pear, qual_match_pear, sizeof t->match_pear);
	memcpy(t->mismatch_pear, qual_mismatch_pear, sizeof t->mismatch_pear);
	memcpy(t->mismatch_rdp, qual_mismatch_rdp_mle, sizeof t->mismatch_rdp);
	memcpy(t->mismatch_rdp_asm, qual_mismatch_assembled_rdp_mle, sizeof t->mismatch_rdp_asm);
	memcpy(t->match_uparse, qual_match_uparse, sizeof t->match_uparse);
	memcpy(t->mismatch_uparse, qual_mismatch_uparse, sizeof t->mismatch_uparse);
	memcpy(t->score, qual_score, sizeof t->score);
	memcpy(t->score_err, qual_score_err, sizeof t->score_err);
}

/* Direct calls into the reference's callbacks / primer scan, for unit parity. */
size_t ref_compute_offset_qual(double threshold, double penalty, int reverse,
                               const po_qual *hay, size_t hay_len, const char *needle, size_t needle_len) {
	return panda_compute_offset_qual(threshold, penalty, reverse != 0, (const panda_qual *) hay, hay_len, (const panda_nt *) needle, needle_len);
}

/* ---- the stages either side of the hot path, through the reference's own entry points ---------- */
typedef struct {
	const char *data;
	size_t len, pos, max_read;
} mem_src;

static bool mem_read(char *buffer, size_t buffer_length, size_t *read, void *user) {
	mem_src *m = user;
	size_t n = m->len - m->pos;
	if (n > buffer_length)
		n = buffer_length;
	if (m->max_read && n > m->max_read)
		n = m->max_read;
	memcpy(buffer, m->data + m->pos, n);
	m->pos += n;
	*read = n;
	return true;
}

typedef struct {
	char text[1 << 16];
	size_t len;
} log_sink;

static void sink_write(const char *buffer, size_t buffer_length, void *user) {
	log_sink *s = user;
	if (s->len + buffer_length >= sizeof s->text)
		buffer_length = sizeof s->text - 1 - s->len;
	memcpy(s->text + s->len, buffer, buffer_length);
	s->len += buffer_length;
	s->text[s->len] = '\0';
}

static int code_from_log(const char *text) {
	/* the LAST "ERR\t<code>" line decides (fastq.c logs exactly one before it stops) */
	static const struct { const char *name; int code; } map[] = {
		{ "ERR\tBADID", PO_FQ_ID_PARSE_FAILURE }, { "ERR\tNOTPAIRED", PO_FQ_NOT_PAIRED }, { "ERR\tEOF", PO_FQ_PREMATURE_EOF },
		{ "ERR\tBADNT", PO_FQ_BAD_NT }, { "ERR\tREADLEN", PO_FQ_READ_TOO_LONG }, { "ERR\tBADSEQ", PO_FQ_PARSE_FAILURE },
		{ "ERR\tNOQUAL", PO_FQ_NO_QUALITY_INFO },
	};
	int code = PO_FQ_OK;
	const char *best = NULL;
	for (size_t k = 0; k < sizeof map / sizeof map[0]; k++) {
		const char *p = text, *last = NULL;
		while ((p = strstr(p, map[k].name)) != NULL) {
			last = p;
			p++;
		}
		if (last != NULL && (best == NULL || last > best)) {
			best = last;
			code = map[k].code;
		}
	}
	return code;
}

/* panda_create_fastq_reader (fastq.c:205-237) over two in-memory files; `max_read` > 0 caps what one
 * PandaBufferRead call returns, to exercise linebuf refills. */
int ref_fastq_parse(const char *fwd, size_t fwd_len, const char *rev, size_t rev_len, int qualmin, int policy,
                    size_t max_pairs, po_fastq_out *out, size_t max_read) {
	mem_src sf = { fwd, fwd_len, 0, max_read }, sr = { rev, rev_len, 0, max_read };
	log_sink *sink = calloc(1, sizeof(log_sink));
	PandaWriter sink_writer = panda_writer_new(sink_write, sink, NULL);
	PandaLogProxy logger = panda_log_proxy_new(sink_writer);
	panda_writer_unref(sink_writer);
	void *next_data = NULL;
	PandaDestroy next_destroy = NULL;
	PandaNextSeq next;
	size_t n = 0;
	uint64_t fo = 0, ro = 0;
	panda_debug_flags = PANDA_DEBUG_FILE;
	next = panda_create_fastq_reader(mem_read, &sf, NULL, mem_read, &sr, NULL, logger, (unsigned char) qualmin, (PandaTagging) policy,
	                                 NULL, NULL, NULL, &next_data, &next_destroy);
	if (out->f_off) out->f_off[0] = 0;
	if (out->r_off) out->r_off[0] = 0;
	while (n < max_pairs) {
		panda_seq_identifier id;
		const panda_qual *f, *r;
		size_t fl, rl;
		memset(&id, 0, sizeof id);
		if (!next(&id, &f, &fl, &r, &rl, next_data))
			break;
		if (out->ids) memcpy(&out->ids[n], &id, sizeof id);
		if (out->f_data) memcpy(out->f_data + fo, f, fl * sizeof(panda_qual));
		if (out->r_data) memcpy(out->r_data + ro, r, rl * sizeof(panda_qual));
		fo += fl;
		ro += rl;
		n++;
		if (out->f_off) out->f_off[n] = fo;
		if (out->r_off) out->r_off[n] = ro;
	}
	out->n = n;
	out->records = 0;      /* not observable through the reference's API */
	/* the writer hands its buffer over when the last reference goes away */
	if (next_destroy)
		next_destroy(next_data);
	panda_log_proxy_unref(logger);
	panda_debug_flags = 0;
	out->error = code_from_log(sink->text);
	free(sink);
	return 0;
}

int ref_seqid_parse(po_seq_identifier *id, const char *input, int policy, int *format) {
	PandaIdFmt fmt = PANDA_IDFMT_UNKNOWN;
	int rc = panda_seqid_parse_fail((panda_seq_identifier *) id, input, (PandaTagging) policy, &fmt, NULL);
	if (format)
		*format = (int) fmt;
	return rc;
}

char ref_result_phred(double p) {
	panda_result r;
	r.nt = 1;
	r.p = p;
	return panda_result_phred(&r);
}

typedef struct {
	char *dst;
	size_t len;
} text_sink;

static void text_write(const char *buffer, size_t buffer_length, void *user) {
	text_sink *s = user;
	memcpy(s->dst + s->len, buffer, buffer_length);
	s->len += buffer_length;
}

/* panda_output_fasta / panda_output_fastq (output.c:85-126) through a PandaWriter into memory */
size_t ref_format_flat(char *dst, int fastq, size_t n, const po_seq_identifier *ids, const uint8_t *status,
                       const double *quality, const int32_t *seq_len, const uint8_t *seq_nt, const double *seq_p,
                       int64_t seq_stride) {
	text_sink sink = { dst, 0 };
	PandaWriter w = panda_writer_new(text_write, &sink, NULL);
	panda_result *seq = calloc(2 * PO_MAX_LEN + 1, sizeof(panda_result));
	for (size_t i = 0; i < n; i++) {
		panda_result_seq rs;
		if (status[i] != PO_OK)
			continue;
		memset(&rs, 0, sizeof rs);
		rs.quality = quality[i];
		memcpy(&rs.name, &ids[i], sizeof rs.name);
		rs.sequence = seq;
		rs.sequence_length = (size_t) seq_len[i];
		for (size_t k = 0; k < rs.sequence_length; k++) {
			seq[k].nt = (panda_nt) seq_nt[i * (size_t) seq_stride + k];
			seq[k].p = seq_p ? seq_p[i * (size_t) seq_stride + k] : 0.0;
		}
		if (fastq)
			panda_output_fastq(&rs, w);
		else
			panda_output_fasta(&rs, w);
	}
	panda_writer_unref(w);
	free(seq);
	return sink.len;
}
