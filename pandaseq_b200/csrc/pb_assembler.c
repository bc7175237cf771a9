/* pb_assembler.c -- the PandaAssembler object over the device batch layer.
 *
 * Same configuration surface, defaults, counters and result ownership rules as
 * the reference's assembler (assembler_support.c:26-410, assembler.c:350-383);
 * the difference is underneath: assemble_seq()/align() are not here.  A call
 * flattens the object into a pb_config, ships the pair(s) to the GPU through
 * pb_assemble_host() and expands the compact device result into the
 * panda_result_seq the caller expects.  panda_assembler_next() and
 * panda_assembler_assemble_batch() amortise that over device-sized batches.
 */
#define _GNU_SOURCE
#include "pb_internal.h"
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEXT_BATCH_DEFAULT 65536	/* pairs pulled from a PandaNextSeq source per launch (env PANDASEQ_B200_NEXT_BATCH overrides) */
#define SEQ_CAP (2 * PB_MAX_LEN + 12)	/* 912: multiple of 16 */

struct panda_assembler {
	volatile size_t refcnt;
	pthread_mutex_t mutex;
	pb_context *ctx;
	PandaLogProxy logger;

	PandaNextSeq next;
	void *next_data;
	PandaDestroy next_destroy;
	PandaFailAlign noalgn;
	void *noalgn_data;
	PandaDestroy noalgn_destroy;

	/* configuration (assembler.h:40-53, 59-78) */
	double threshold;	/* log */
	size_t minoverlap, maxoverlap, num_kmers;
	PandaAlgorithm algo;
	size_t forward_primer_length, reverse_primer_length, forward_trim, reverse_trim;
	bool post_primers;
	double primer_penalty;
	panda_nt forward_primer[PB_MAX_LEN], reverse_primer[PB_MAX_LEN];
	char name[PB_MAX_LEN];

	/* modules with host callbacks (assembler.h:69-73, module.c:156-180) */
	PandaModule *modules;
	size_t *rejected;
	size_t modules_length, modules_size;
	bool has_check;		/* some module checks assembled pairs: OK / overlap counters are then kept on the host, after the checks */

	/* counters */
	int64_t counters[PB_NCOUNTERS];

	/* the single result the API hands out (assembler.h:55, 74) */
	panda_result_seq result;
	panda_result result_seq[SEQ_CAP];

	/* batch state for panda_assembler_next / assemble_batch */
	size_t cap_pairs, cap_f, cap_r;
	panda_seq_identifier *ids;
	panda_qual *f_data, *r_data;
	uint64_t *f_off, *r_off;
	pb_pair_result *res;
	uint8_t *nt;
	double *p;
	size_t batch_n, batch_pos, next_batch;
	bool source_dry;
};

static void flatten(PandaAssembler a, pb_config *cfg, bool *ok) {
	pb_config_default(cfg, PB_SIMPLE_BAYES);
	*ok = pb_algorithm_fill_config(a->algo, cfg) == 0;
	cfg->post_primers = a->post_primers;
	cfg->minoverlap = (int64_t) a->minoverlap;
	cfg->maxoverlap = (int64_t) a->maxoverlap;
	cfg->num_kmers = (int64_t) a->num_kmers;
	cfg->forward_trim = (int64_t) a->forward_trim;
	cfg->reverse_trim = (int64_t) a->reverse_trim;
	cfg->forward_primer_length = (int64_t) a->forward_primer_length;
	cfg->reverse_primer_length = (int64_t) a->reverse_primer_length;
	cfg->threshold = a->threshold;
	cfg->primer_penalty = a->primer_penalty;
	memcpy(cfg->forward_primer, a->forward_primer, a->forward_primer_length);
	memcpy(cfg->reverse_primer, a->reverse_primer, a->reverse_primer_length);
}

PandaAssembler panda_assembler_new(PandaNextSeq next, void *next_data, PandaDestroy next_destroy, PandaLogProxy logger) {
	return panda_assembler_new_kmer(next, next_data, next_destroy, logger, PANDA_DEFAULT_NUM_KMERS);
}

PandaAssembler panda_assembler_new_kmer(PandaNextSeq next, void *next_data, PandaDestroy next_destroy, PandaLogProxy logger, size_t num_kmers) {
	pb_context *ctx = NULL;
	PandaAssembler a = NULL;
	if (num_kmers != 2) {
		pb_set_error("num_kmers must be 2 on the device path");
	} else if (pb_shared_context(&ctx) == PB_OK) {
		a = calloc(1, sizeof(struct panda_assembler));
	}
	if (a == NULL) {
		if (next_destroy != NULL)
			next_destroy(next_data);
		return NULL;
	}
	pthread_mutex_init(&a->mutex, NULL);
	a->refcnt = 1;
	a->ctx = ctx;
	a->logger = logger;
	a->next = next;
	a->next_data = next_data;
	a->next_destroy = next_destroy;
	a->threshold = log(0.6);
	a->minoverlap = 2;
	a->maxoverlap = 0;
	a->num_kmers = num_kmers;
	a->algo = panda_algorithm_simple_bayes_new();
	a->result.sequence = a->result_seq;
	a->next_batch = NEXT_BATCH_DEFAULT;
	{
		const char *env = getenv("PANDASEQ_B200_NEXT_BATCH");
		long v = env ? atol(env) : 0;
		if (v > 0 && v <= (1L << 24))
			a->next_batch = (size_t) v;
	}
	return a;
}

/* ---- modules (module.c:31-98, 124-216; pandaseq-module.h) ------------------------------------------ */

struct panda_module {
	volatile size_t refcnt;
	pthread_mutex_t mutex;
	char *name;
	PandaCheck check;
	PandaPreCheck precheck;
	void *user_data;
	PandaDestroy cleanup;
};

PandaModule panda_module_new(const char *name, PandaCheck check, PandaPreCheck precheck, void *user_data, PandaDestroy cleanup) {
	PandaModule m;
	if (name == NULL || (check == NULL && precheck == NULL))	/* module.c:258-260 */
		return NULL;
	m = calloc(1, sizeof *m);
	if (m == NULL)
		return NULL;
	pthread_mutex_init(&m->mutex, NULL);
	m->refcnt = 1;
	m->name = malloc(strlen(name) + 1);
	if (m->name == NULL) {
		free(m);
		return NULL;
	}
	strcpy(m->name, name);
	m->check = check;
	m->precheck = precheck;
	m->user_data = user_data;
	m->cleanup = cleanup;
	return m;
}

PandaModule panda_module_ref(PandaModule m) {
	pthread_mutex_lock(&m->mutex);
	m->refcnt++;
	pthread_mutex_unlock(&m->mutex);
	return m;
}

void panda_module_unref(PandaModule m) {
	size_t left;
	if (m == NULL)
		return;
	pthread_mutex_lock(&m->mutex);
	left = --m->refcnt;
	pthread_mutex_unlock(&m->mutex);
	if (left != 0)
		return;
	pthread_mutex_destroy(&m->mutex);
	if (m->cleanup != NULL)
		m->cleanup(m->user_data);
	free(m->name);
	free(m);
}

const char *panda_module_get_name(PandaModule m) { return m->name; }
int panda_module_get_api(PandaModule m) { (void) m; return 3; }	/* PANDA_API, pandaseq.h:61 */

bool panda_assembler_add_module(PandaAssembler a, PandaModule m) {
	if (m == NULL)
		return false;
	pthread_mutex_lock(&a->mutex);
	if (a->modules_length == a->modules_size) {
		size_t size = a->modules_size == 0 ? 8 : a->modules_size * 2;
		PandaModule *mods = realloc(a->modules, size * sizeof(PandaModule));
		size_t *rej = mods ? realloc(a->rejected, size * sizeof(size_t)) : NULL;
		if (mods)
			a->modules = mods;
		if (rej)
			a->rejected = rej;
		if (!mods || !rej) {
			pthread_mutex_unlock(&a->mutex);
			return false;
		}
		a->modules_size = size;
	}
	a->rejected[a->modules_length] = 0;
	a->modules[a->modules_length++] = panda_module_ref(m);
	if (m->check != NULL)
		a->has_check = true;
	pthread_mutex_unlock(&a->mutex);
	return true;
}

size_t panda_assembler_add_modules(PandaAssembler a, PandaModule *modules, size_t modules_length) {
	size_t it;
	for (it = 0; it < modules_length; it++)
		if (!panda_assembler_add_module(a, modules[it]))
			return it;
	return it;
}

bool panda_assembler_foreach_module(PandaAssembler a, PandaModuleCallback callback, void *data) {
	for (size_t it = 0; it < a->modules_length; it++)
		if (!callback(a, a->modules[it], a->rejected[it], data))
			return false;
	return true;
}

void panda_assembler_module_stats(PandaAssembler a) { (void) a; }	/* module.c:208-216 only logs */

/* module_precheckseq (module.c:139-154) where assemble_seq calls it (assembler.c:255-261): after the pair was counted and
 * found to have two bases per read.  A rejected pair is counted here and never staged for the device. */
static bool host_precheck(PandaAssembler a, const panda_seq_identifier *id, const panda_qual *f, size_t fl, const panda_qual *r, size_t rl) {
	static const panda_seq_identifier no_id;
	if (a->modules_length == 0 || fl < 2 || rl < 2)
		return true;
	for (size_t it = 0; it < a->modules_length; it++) {
		PandaModule m = a->modules[it];
		if (m->precheck != NULL && !m->precheck(a->logger, id ? id : &no_id, f, fl, r, rl, m->user_data)) {
			a->rejected[it]++;
			a->counters[PB_C_COUNT]++;
			return false;
		}
	}
	return true;
}

/* module_checkseq (module.c:124-137) and what assemble_seq does when it passes (assembler.c:339-346) */
static bool host_check(PandaAssembler a, const panda_result_seq *res) {
	if (!a->has_check)
		return true;
	for (size_t it = 0; it < a->modules_length; it++) {
		PandaModule m = a->modules[it];
		if (m->check != NULL && !m->check(a->logger, res, m->user_data)) {
			a->rejected[it]++;
			return false;
		}
	}
	a->counters[PB_C_OK]++;
	a->counters[PB_C_OVERLAPS + res->overlap]++;
	if (a->counters[PB_C_LONGEST] < (int64_t) res->overlap)
		a->counters[PB_C_LONGEST] = (int64_t) res->overlap;
	return true;
}

PandaAssembler panda_assembler_ref(PandaAssembler a) {
	pthread_mutex_lock(&a->mutex);
	a->refcnt++;
	pthread_mutex_unlock(&a->mutex);
	return a;
}

void panda_assembler_unref(PandaAssembler a) {
	size_t left;
	if (a == NULL)
		return;
	pthread_mutex_lock(&a->mutex);
	left = --a->refcnt;
	pthread_mutex_unlock(&a->mutex);
	if (left != 0)
		return;
	pthread_mutex_destroy(&a->mutex);
	if (a->next_destroy != NULL && a->next != NULL)
		a->next_destroy(a->next_data);
	if (a->noalgn_destroy != NULL && a->noalgn != NULL)
		a->noalgn_destroy(a->noalgn_data);
	panda_algorithm_unref(a->algo);
	for (size_t it = 0; it < a->modules_length; it++)
		panda_module_unref(a->modules[it]);
	free(a->modules);
	free(a->rejected);
	free(a->ids);
	free(a->f_data);
	free(a->r_data);
	free(a->f_off);
	free(a->r_off);
	free(a->res);
	free(a->nt);
	free(a->p);
	free(a);
}

void panda_assembler_copy_configuration(PandaAssembler dest, PandaAssembler src) {
	for (size_t it = 0; it < src->modules_length; it++)	/* assembler_support.c:123-125 */
		panda_assembler_add_module(dest, src->modules[it]);
	panda_assembler_set_forward_primer(dest, src->forward_primer, src->forward_primer_length);
	panda_assembler_set_reverse_primer(dest, src->reverse_primer, src->reverse_primer_length);
	dest->forward_trim = src->forward_trim;
	dest->reverse_trim = src->reverse_trim;
	dest->threshold = src->threshold;
	dest->minoverlap = src->minoverlap;
	dest->maxoverlap = src->maxoverlap;
	dest->post_primers = src->post_primers;
	panda_algorithm_unref(dest->algo);
	dest->algo = panda_algorithm_ref(src->algo);
	dest->primer_penalty = src->primer_penalty;
}

/* ---- getters / setters: same guards as the reference's ---------------------------- */

PandaAlgorithm panda_assembler_get_algorithm(PandaAssembler a) { return a->algo; }
void panda_assembler_set_algorithm(PandaAssembler a, PandaAlgorithm algorithm) {
	if (algorithm == NULL)
		return;
	panda_algorithm_unref(a->algo);
	a->algo = panda_algorithm_ref(algorithm);
}
long panda_assembler_get_bad_read_count(PandaAssembler a) { return (long) a->counters[PB_C_BADR]; }
long panda_assembler_get_count(PandaAssembler a) { return (long) a->counters[PB_C_COUNT]; }
long panda_assembler_get_failed_alignment_count(PandaAssembler a) { return (long) a->counters[PB_C_NOALGN]; }
long panda_assembler_get_low_quality_count(PandaAssembler a) { return (long) a->counters[PB_C_LOWQ]; }
long panda_assembler_get_no_forward_primer_count(PandaAssembler a) { return (long) a->counters[PB_C_NOFP]; }
long panda_assembler_get_no_reverse_primer_count(PandaAssembler a) { return (long) a->counters[PB_C_NORP]; }
long panda_assembler_get_ok_count(PandaAssembler a) { return (long) a->counters[PB_C_OK]; }
long panda_assembler_get_slow_count(PandaAssembler a) { return (long) a->counters[PB_C_SLOW]; }
size_t panda_assembler_get_longest_overlap(PandaAssembler a) { return (size_t) a->counters[PB_C_LONGEST]; }
long panda_assembler_get_overlap_count(PandaAssembler a, size_t overlap) {
	return overlap < 2 * PB_MAX_LEN ? (long) a->counters[PB_C_OVERLAPS + overlap] : -1;
}
size_t panda_assembler_get_num_kmer(PandaAssembler a) { return a->num_kmers; }
PandaLogProxy panda_assembler_get_logger(PandaAssembler a) { return a->logger; }

void panda_assembler_set_fail_alignment(PandaAssembler a, PandaFailAlign handler, void *handler_data, PandaDestroy handler_destroy) {
	if (a->noalgn_destroy != NULL && a->noalgn != NULL)
		a->noalgn_destroy(a->noalgn_data);
	a->noalgn = handler;
	a->noalgn_data = handler_data;
	a->noalgn_destroy = handler_destroy;
}

panda_nt *panda_assembler_get_forward_primer(PandaAssembler a, size_t *length) {
	if (length != NULL)
		*length = a->forward_primer_length;
	return a->forward_primer_length == 0 ? NULL : a->forward_primer;
}
void panda_assembler_set_forward_primer(PandaAssembler a, panda_nt *sequence, size_t length) {
	if (length >= PB_MAX_LEN)
		return;
	memcpy(a->forward_primer, sequence, length);
	a->forward_primer_length = length;
	a->forward_trim = 0;
}
size_t panda_assembler_get_forward_trim(PandaAssembler a) { return a->forward_trim; }
void panda_assembler_set_forward_trim(PandaAssembler a, size_t trim) {
	a->forward_trim = trim;
	a->forward_primer_length = 0;
}
panda_nt *panda_assembler_get_reverse_primer(PandaAssembler a, size_t *length) {
	if (length != NULL)
		*length = a->reverse_primer_length;
	return a->reverse_primer_length == 0 ? NULL : a->reverse_primer;
}
void panda_assembler_set_reverse_primer(PandaAssembler a, panda_nt *sequence, size_t length) {
	if (length >= PB_MAX_LEN)
		return;
	memcpy(a->reverse_primer, sequence, length);
	a->reverse_primer_length = length;
	a->reverse_trim = 0;
}
size_t panda_assembler_get_reverse_trim(PandaAssembler a) { return a->reverse_trim; }
void panda_assembler_set_reverse_trim(PandaAssembler a, size_t trim) {
	a->reverse_trim = trim;
	a->reverse_primer_length = 0;
}
int panda_assembler_get_minimum_overlap(PandaAssembler a) { return (int) a->minoverlap; }
void panda_assembler_set_minimum_overlap(PandaAssembler a, int overlap) {
	if (overlap > 1 && (size_t) overlap < 2 * PB_MAX_LEN)
		a->minoverlap = (size_t) overlap;
}
int panda_assembler_get_maximum_overlap(PandaAssembler a) { return (int) a->maxoverlap; }
void panda_assembler_set_maximum_overlap(PandaAssembler a, int overlap) {
	if (overlap >= 0 && (size_t) overlap < 2 * PB_MAX_LEN)
		a->maxoverlap = (size_t) overlap;
}
const char *panda_assembler_get_name(PandaAssembler a) {
	return (a == NULL || a->name[0] == '\0') ? NULL : a->name;
}
void panda_assembler_set_name(PandaAssembler a, const char *name) {
	if (name == NULL) {
		a->name[0] = '\0';
		return;
	}
	strncpy(a->name, name, PB_MAX_LEN);
	a->name[PB_MAX_LEN - 1] = '\0';
}
bool panda_assembler_get_primers_after(PandaAssembler a) { return a->post_primers; }
void panda_assembler_set_primers_after(PandaAssembler a, bool after) { a->post_primers = after; }
double panda_assembler_get_threshold(PandaAssembler a) { return exp(a->threshold); }
void panda_assembler_set_threshold(PandaAssembler a, double threshold) {
	if (threshold > 0 && threshold < 1)
		a->threshold = log(threshold);
}
double panda_assembler_get_primer_penalty(PandaAssembler a) { return exp(a->primer_penalty); }	/* sic: assembler_support.c:399-402 */
void panda_assembler_set_primer_penalty(PandaAssembler a, double penalty) {
	if (penalty >= 0 && penalty < 1)
		a->primer_penalty = penalty;
}

/* ---- batches ------------------------------------------------------------------------ */

static bool reserve(PandaAssembler a, size_t pairs, size_t fbases, size_t rbases) {
	if (pairs > a->cap_pairs) {
		size_t cap = pairs < 64 ? 64 : pairs;
		void *ids = realloc(a->ids, cap * sizeof(panda_seq_identifier));
		if (ids) a->ids = ids;
		void *fo = realloc(a->f_off, (cap + 1) * sizeof(uint64_t));
		if (fo) a->f_off = fo;
		void *ro = realloc(a->r_off, (cap + 1) * sizeof(uint64_t));
		if (ro) a->r_off = ro;
		void *res = realloc(a->res, cap * sizeof(pb_pair_result));
		if (res) a->res = res;
		void *nt = realloc(a->nt, cap * (SEQ_CAP / 2));
		if (nt) a->nt = nt;
		void *p = realloc(a->p, cap * SEQ_CAP * sizeof(double));
		if (p) a->p = p;
		if (!ids || !fo || !ro || !res || !nt || !p)
			return false;
		a->cap_pairs = cap;
	}
	if (fbases > a->cap_f) {
		size_t cap = fbases + fbases / 2 + 1024;
		void *f = realloc(a->f_data, cap * sizeof(panda_qual));
		if (!f)
			return false;
		a->f_data = f;
		a->cap_f = cap;
	}
	if (rbases > a->cap_r) {
		size_t cap = rbases + rbases / 2 + 1024;
		void *r = realloc(a->r_data, cap * sizeof(panda_qual));
		if (!r)
			return false;
		a->r_data = r;
		a->cap_r = cap;
	}
	return true;
}

/* Run the staged batch [0, a->batch_n) on the device and fold its counters in. */
static bool run_batch(PandaAssembler a) {
	pb_config cfg;
	bool ok;
	flatten(a, &cfg, &ok);
	if (!ok)
		return false;
	if (!a->has_check)
		return pb_assemble_host(a->ctx, &cfg, a->batch_n, a->f_data, a->f_off, a->r_data, a->r_off,
		                        a->res, a->nt, a->p, SEQ_CAP, a->counters) == PB_OK;
	/* a module checks the assembled pairs on the host: the device's count of accepted pairs (and their overlap histogram) is
	 * taken before those checks, so these counters are kept by host_check() as the pairs are handed out */
	int64_t tmp[PB_NCOUNTERS];
	memset(tmp, 0, sizeof tmp);
	if (pb_assemble_host(a->ctx, &cfg, a->batch_n, a->f_data, a->f_off, a->r_data, a->r_off,
	                     a->res, a->nt, a->p, SEQ_CAP, tmp) != PB_OK)
		return false;
	tmp[PB_C_OK] = 0;
	tmp[PB_C_LONGEST] = 0;
	memset(tmp + PB_C_OVERLAPS, 0, (PB_NCOUNTERS - PB_C_OVERLAPS) * sizeof(int64_t));
	pb_counters_merge(a->counters, tmp);
	return true;
}

/* Expand device record i of the staged batch into a->result (pandaseq-common.h:277-330). */
static const panda_result_seq *publish(PandaAssembler a, size_t i, const panda_seq_identifier *id,
                                       const panda_qual *fwd, size_t flen, const panda_qual *rev, size_t rlen) {
	const pb_pair_result *r = &a->res[i];
	panda_result_seq *out = &a->result;
	if (id != NULL)
		out->name = *id;
	out->forward = fwd;
	out->forward_length = flen;
	out->reverse = rev;
	out->reverse_length = rlen;
	out->sequence = a->result_seq;
	out->quality = r->quality;
	out->degenerates = r->degenerates;
	out->sequence_length = r->seq_len;
	out->forward_offset = r->fwd_offset;
	out->reverse_offset = r->rev_offset;
	out->overlap_mismatches = r->mismatches;
	out->overlaps_examined = r->examined;
	out->overlap = r->overlap;
	out->estimated_overlap_probability = r->est_prob;
	const uint8_t *nt = a->nt + i * (SEQ_CAP / 2);	/* 4 bit per base, base 2k in the low nibble */
	const double *p = a->p + i * SEQ_CAP;
	for (size_t k = 0; k < r->seq_len; k++) {
		a->result_seq[k].nt = (panda_nt) ((nt[k >> 1] >> ((k & 1) * 4)) & 0x0F);
		a->result_seq[k].p = p[k];
	}
	return out;
}

const panda_result_seq *panda_assembler_assemble(PandaAssembler a, panda_seq_identifier *id,
                                                 const panda_qual *forward, size_t forward_length,
                                                 const panda_qual *reverse, size_t reverse_length) {
	assert(forward_length <= PB_MAX_LEN);
	assert(reverse_length <= PB_MAX_LEN);
	if (!host_precheck(a, id, forward, forward_length, reverse, reverse_length))
		return NULL;
	if (!reserve(a, 1, forward_length, reverse_length))
		return NULL;
	memcpy(a->f_data, forward, forward_length * sizeof(panda_qual));
	memcpy(a->r_data, reverse, reverse_length * sizeof(panda_qual));
	a->f_off[0] = a->r_off[0] = 0;
	a->f_off[1] = forward_length;
	a->r_off[1] = reverse_length;
	a->batch_n = 1;
	a->batch_pos = 1;	/* not part of a next() stream */
	if (!run_batch(a))
		return NULL;
	if (a->res[0].status == PB_PAIR_NOALGN && a->noalgn != NULL)
		a->noalgn(a, id, forward, forward_length, reverse, reverse_length, a->noalgn_data);
	if (a->res[0].status != PB_PAIR_OK)
		return NULL;
	const panda_result_seq *out = publish(a, 0, id, forward, forward_length, reverse, reverse_length);
	return host_check(a, out) ? out : NULL;
}

size_t panda_assembler_assemble_batch(PandaAssembler a, size_t n, const panda_seq_identifier *ids,
                                      const panda_qual *const *forward, const size_t *forward_length,
                                      const panda_qual *const *reverse, const size_t *reverse_length,
                                      PandaOutputSeq output, void *output_data) {
	size_t fb = 0, rb = 0, accepted = 0;
	for (size_t i = 0; i < n; i++) {
		if (forward_length[i] > PB_MAX_LEN || reverse_length[i] > PB_MAX_LEN) {
			pb_set_error("pair %zu longer than PANDA_MAX_LEN", i);
			return (size_t) -1;
		}
		fb += forward_length[i];
		rb += reverse_length[i];
	}
	if (!reserve(a, n, fb, rb))
		return (size_t) -1;
	/* pairs a module's pre-check rejects are not staged; map[k] = caller's index of staged pair k */
	size_t *map = a->modules_length ? malloc((n ? n : 1) * sizeof(size_t)) : NULL;
	if (a->modules_length && map == NULL)
		return (size_t) -1;
	size_t m = 0;
	fb = rb = 0;
	for (size_t i = 0; i < n; i++) {
		if (!host_precheck(a, ids ? &ids[i] : NULL, forward[i], forward_length[i], reverse[i], reverse_length[i]))
			continue;
		if (map)
			map[m] = i;
		a->f_off[m] = fb;
		a->r_off[m] = rb;
		memcpy(a->f_data + fb, forward[i], forward_length[i] * sizeof(panda_qual));
		memcpy(a->r_data + rb, reverse[i], reverse_length[i] * sizeof(panda_qual));
		fb += forward_length[i];
		rb += reverse_length[i];
		m++;
	}
	a->f_off[m] = fb;
	a->r_off[m] = rb;
	a->batch_n = m;
	a->batch_pos = m;
	if (m > 0 && !run_batch(a)) {
		free(map);
		return (size_t) -1;
	}
	for (size_t k = 0; k < m; k++) {
		const size_t i = map ? map[k] : k;
		const panda_seq_identifier *id = ids ? &ids[i] : NULL;
		if (a->res[k].status == PB_PAIR_NOALGN && a->noalgn != NULL)
			a->noalgn(a, id, forward[i], forward_length[i], reverse[i], reverse_length[i], a->noalgn_data);
		if (a->res[k].status != PB_PAIR_OK)
			continue;
		const panda_result_seq *out = publish(a, k, id, forward[i], forward_length[i], reverse[i], reverse_length[i]);
		if (!host_check(a, out))
			continue;
		accepted++;
		if (output != NULL)
			output(out, output_data);
	}
	free(map);
	return accepted;
}

const panda_result_seq *panda_assembler_next(PandaAssembler a) {
	if (a->next == NULL)
		return NULL;
	for (;;) {
		/* hand out what the last launch produced, in input order */
		while (a->batch_pos < a->batch_n) {
			size_t i = a->batch_pos++;
			const panda_qual *f = a->f_data + a->f_off[i], *r = a->r_data + a->r_off[i];
			size_t fl = (size_t) (a->f_off[i + 1] - a->f_off[i]), rl = (size_t) (a->r_off[i + 1] - a->r_off[i]);
			if (a->res[i].status == PB_PAIR_NOALGN && a->noalgn != NULL)
				a->noalgn(a, &a->ids[i], f, fl, r, rl, a->noalgn_data);
			if (a->res[i].status == PB_PAIR_OK) {
				const panda_result_seq *out = publish(a, i, &a->ids[i], f, fl, r, rl);
				if (host_check(a, out))
					return out;
			}
		}
		if (a->source_dry)
			return NULL;
		/* refill: the source's arrays are only valid until its next call, so copy as we pull
		 * (the reference's mux does the same, mux.c:150-157) */
		size_t n = 0, fb = 0, rb = 0;
		if (!reserve(a, a->next_batch, a->next_batch * 160, a->next_batch * 160))
			return NULL;
		while (n < a->next_batch) {
			const panda_qual *f, *r;
			size_t fl, rl;
			if (!a->next(&a->ids[n], &f, &fl, &r, &rl, a->next_data)) {
				a->source_dry = true;
				break;
			}
			assert(fl <= PB_MAX_LEN);
			assert(rl <= PB_MAX_LEN);
			if (!host_precheck(a, &a->ids[n], f, fl, r, rl))
				continue;	/* counted and rejected on the host: not staged */
			if (!reserve(a, n + 1, fb + fl, rb + rl))
				return NULL;
			a->f_off[n] = fb;
			a->r_off[n] = rb;
			memcpy(a->f_data + fb, f, fl * sizeof(panda_qual));
			memcpy(a->r_data + rb, r, rl * sizeof(panda_qual));
			fb += fl;
			rb += rl;
			n++;
		}
		a->f_off[n] = fb;
		a->r_off[n] = rb;
		a->batch_n = n;
		a->batch_pos = 0;
		if (n == 0)
			return NULL;
		if (!run_batch(a)) {
			a->batch_n = a->batch_pos = 0;
			return NULL;
		}
	}
}

/* pool.c:110-181.  The reference fans one assembler out over `threads` workers that pull from a shared PandaMux; here
 * the one assembler already drains its source in device-sized batches (panda_assembler_next), so `threads` and `mux` only
 * keep the signature: the mux is an opaque handle this library never creates, and is ignored.  Ownership follows the
 * reference: the assembler is consumed (unref), output_destroy(output_data) is called at the end.  Returns whether any
 * pair was read. */
bool panda_run_pool(int threads, PandaAssembler assembler, PandaMux mux, PandaOutputSeq output, void *output_data, PandaDestroy output_destroy) {
	const panda_result_seq *result;
	bool some_seqs;
	(void) threads;
	(void) mux;
	if (assembler == NULL) {
		if (output_destroy != NULL)
			output_destroy(output_data);
		return false;
	}
	while ((result = panda_assembler_next(assembler)) != NULL) {
		if (output != NULL && !output(result, output_data))
			break;
	}
	some_seqs = panda_assembler_get_count(assembler) > 0;
	panda_assembler_unref(assembler);
	if (output_destroy != NULL)
		output_destroy(output_data);
	return some_seqs;
}

/* offset.c:103-112 as a batch of one read: an assembler-free entry point.  The read is
 * paired with a 2-base dummy mate and run through the kernel with the needle as forward
 * primer; the forward offset the kernel reports is bestindex-1. */
size_t panda_compute_offset_qual(double threshold, double penalty, bool reverse,
                                 const panda_qual *haystack, size_t haystack_length,
                                 const panda_nt *needle, size_t needle_length) {
	pb_context *ctx;
	pb_config cfg;
	pb_pair_result res;
	panda_qual *hay;
	panda_qual mate[2] = { { PANDA_NT_A, 30 }, { PANDA_NT_A, 30 } };
	uint64_t f_off[2] = { 0, haystack_length }, r_off[2] = { 0, 2 };
	if (needle_length == 0 || needle_length >= PB_MAX_LEN || haystack_length > PB_MAX_LEN || haystack_length < 2)
		return 0;
	if (pb_shared_context(&ctx) != PB_OK)
		return 0;
	hay = malloc(haystack_length * sizeof(panda_qual));
	if (hay == NULL)
		return 0;
	for (size_t i = 0; i < haystack_length; i++)
		hay[i] = haystack[reverse ? haystack_length - 1 - i : i];	/* offset.c:79: reverse scans from the end */
	pb_config_default(&cfg, PB_SIMPLE_BAYES);
	cfg.threshold = threshold;
	cfg.primer_penalty = penalty;
	cfg.forward_primer_length = (int64_t) needle_length;
	memcpy(cfg.forward_primer, needle, needle_length);
	pb_status st = pb_assemble_host(ctx, &cfg, 1, hay, f_off, mate, r_off, &res, NULL, NULL, 0, NULL);
	free(hay);
	if (st != PB_OK || res.status == PB_PAIR_NOFP)
		return 0;
	return (size_t) res.fwd_offset + 1;
}
