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

/* One batch on its way through the device.  The arrays the device path copies from / into are page-locked (pb_host_alloc),
 * so pb_assemble_host moves them without a staging copy; the per-base log p comes back as 16-bit codes into the posterior
 * table (pb_assemble_host_codes) and rows are as long as the batch's longest F + R, not 912 doubles. */
struct pb_stage {
	size_t cap_pairs, cap_f, cap_r, cap_res, cap_nt, cap_code, cap_p;
	panda_seq_identifier *ids;
	panda_qual *f_data, *r_data;
	uint64_t *f_off, *r_off;
	pb_pair_result *res;
	uint8_t *nt;
	uint16_t *code;
	double *p;		/* instead of code when the configuration has no codes (primers-after) */
	bool has_codes;
	size_t n, pos, stride, max_seq;
};

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

	/* batch state: the stream of panda_assembler_next() and the calls (panda_assembler_assemble / _assemble_batch) have a
	 * stage each, so that a call between two next()s does not disturb the batch next() is handing out (assembler.c:350-383:
	 * the two are independent in the reference) */
	struct pb_stage stream, call;
	size_t next_batch;
	bool source_dry;
	bool failed;		/* a device call failed: next() stays at the end of its stream, pb_last_error() has the reason */
	/* the posterior table the codes index (pb_posterior_table), rebuilt when the configuration changes */
	double *ptable;
	pb_config ptable_cfg;
	bool ptable_valid;
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

/* Page-locked memory is expensive to get (tens of milliseconds for a stage of 64 K pairs), and panda_run_pool makes a clone
 * with its own stage per worker and per call: stages that are given up go to a small process-wide free list and are handed
 * to the next assembler that needs one. */
#define STAGE_POOL 64
static struct pb_stage stage_pool[STAGE_POOL];
static int stage_pool_n;
static pthread_mutex_t stage_pool_lock = PTHREAD_MUTEX_INITIALIZER;

static void stage_free(struct pb_stage *st) {
	bool kept = false;
	if (st->cap_pairs == 0 && st->ids == NULL)
		return;
	pthread_mutex_lock(&stage_pool_lock);
	if (stage_pool_n < STAGE_POOL) {
		stage_pool[stage_pool_n++] = *st;
		kept = true;
	}
	pthread_mutex_unlock(&stage_pool_lock);
	if (!kept) {
		free(st->ids);
		pb_host_free(st->f_data);
		pb_host_free(st->r_data);
		pb_host_free(st->f_off);
		pb_host_free(st->r_off);
		pb_host_free(st->res);
		pb_host_free(st->nt);
		pb_host_free(st->code);
		pb_host_free(st->p);
	}
	memset(st, 0, sizeof *st);
}

/* an empty stage takes over the largest pooled one, if there is any */
static void stage_adopt(struct pb_stage *st) {
	if (st->cap_pairs != 0 || st->ids != NULL)
		return;
	pthread_mutex_lock(&stage_pool_lock);
	int best = -1;
	for (int k = 0; k < stage_pool_n; k++)
		if (best < 0 || stage_pool[k].cap_pairs > stage_pool[best].cap_pairs)
			best = k;
	if (best >= 0) {
		*st = stage_pool[best];
		stage_pool[best] = stage_pool[--stage_pool_n];
	}
	pthread_mutex_unlock(&stage_pool_lock);
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
	stage_free(&a->stream);
	stage_free(&a->call);
	free(a->ptable);
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

/* grow a page-locked array, keeping `keep` bytes of its content; false (array untouched) when out of memory */
static bool grow_pinned(void **p, size_t *cap, size_t need, size_t elem, size_t keep) {
	if (need <= *cap)
		return true;
	void *q = pb_host_alloc(need * elem);
	if (q == NULL)
		return false;
	if (*p != NULL && keep > 0)
		memcpy(q, *p, keep);
	pb_host_free(*p);
	*p = q;
	*cap = need;
	return true;
}

/* room for `pairs` pairs with `fbases` / `rbases` bases in a stage's input arrays; `used_*` = what is already staged */
static bool reserve(struct pb_stage *st, size_t pairs, size_t fbases, size_t rbases, size_t used_pairs, size_t used_f, size_t used_r) {
	if (pairs > st->cap_pairs) {
		size_t cap = pairs < 64 ? 64 : pairs + pairs / 2, c1 = st->cap_pairs ? st->cap_pairs + 1 : 0, c2 = c1;
		void *ids = realloc(st->ids, cap * sizeof(panda_seq_identifier));
		if (ids == NULL)
			return false;
		st->ids = ids;
		if (!grow_pinned((void **) &st->f_off, &c1, cap + 1, sizeof(uint64_t), (used_pairs + 1) * sizeof(uint64_t))
		    || !grow_pinned((void **) &st->r_off, &c2, cap + 1, sizeof(uint64_t), (used_pairs + 1) * sizeof(uint64_t)))
			return false;		/* cap_pairs unchanged: the arrays that did grow are simply larger than recorded */
		st->cap_pairs = cap;
	}
	if (fbases > st->cap_f && !grow_pinned((void **) &st->f_data, &st->cap_f, fbases + fbases / 2 + 1024, sizeof(panda_qual), used_f * sizeof(panda_qual)))
		return false;
	if (rbases > st->cap_r && !grow_pinned((void **) &st->r_data, &st->cap_r, rbases + rbases / 2 + 1024, sizeof(panda_qual), used_r * sizeof(panda_qual)))
		return false;
	return true;
}

static void stage_begin(struct pb_stage *st) {
	st->n = st->pos = 0;
	st->max_seq = 0;
}

/* copy one pair behind the staged ones (the source's arrays are only valid until its next call: mux.c:150-157 copies too) */
static bool stage_push(struct pb_stage *st, const panda_seq_identifier *id, const panda_qual *f, size_t fl, const panda_qual *r, size_t rl) {
	const size_t n = st->n, fb = n ? (size_t) st->f_off[n] : 0, rb = n ? (size_t) st->r_off[n] : 0;
	if (!reserve(st, n + 1, fb + fl, rb + rl, n, fb, rb))
		return false;
	if (id != NULL)
		st->ids[n] = *id;
	else
		memset(&st->ids[n], 0, sizeof st->ids[n]);
	st->f_off[n] = fb;
	st->r_off[n] = rb;
	memcpy(st->f_data + fb, f, fl * sizeof(panda_qual));
	memcpy(st->r_data + rb, r, rl * sizeof(panda_qual));
	st->f_off[n + 1] = fb + fl;
	st->r_off[n + 1] = rb + rl;
	if (fl + rl > st->max_seq)
		st->max_seq = fl + rl;
	st->n = n + 1;
	return true;
}

/* Run the staged batch [0, st->n) on the device and fold its counters in. */
static bool run_batch(PandaAssembler a, struct pb_stage *st) {
	pb_config cfg;
	bool ok;
	int64_t tmp[PB_NCOUNTERS];
	int64_t *cnt = a->has_check ? tmp : a->counters;
	pb_status rc;
	st->pos = 0;
	if (st->n == 0)
		return true;
	flatten(a, &cfg, &ok);
	if (!ok)
		return false;
	st->stride = (st->max_seq + 15) & ~(size_t) 15;
	if (st->stride < 16)
		st->stride = 16;
	st->has_codes = !cfg.post_primers;	/* the staged sequence of primers-after exists as doubles only (pb_assemble_host_codes) */
	if (!grow_pinned((void **) &st->res, &st->cap_res, st->n, sizeof(pb_pair_result), 0)
	    || !grow_pinned((void **) &st->nt, &st->cap_nt, st->n * (st->stride / 2), 1, 0))
		return false;
	if (st->has_codes) {
		if (!a->ptable_valid || memcmp(&a->ptable_cfg, &cfg, sizeof cfg) != 0) {
			if (a->ptable == NULL)
				a->ptable = malloc(PB_POSTERIOR_CODES * sizeof(double));
			if (a->ptable == NULL || pb_posterior_table(&cfg, a->ptable) != PB_OK)
				return false;
			a->ptable_cfg = cfg;
			a->ptable_valid = true;
		}
		if (!grow_pinned((void **) &st->code, &st->cap_code, st->n * st->stride, sizeof(uint16_t), 0))
			return false;
	} else if (!grow_pinned((void **) &st->p, &st->cap_p, st->n * st->stride, sizeof(double), 0)) {
		return false;
	}
	/* when a module checks the assembled pairs on the host, the device's count of accepted pairs (and their overlap
	 * histogram) is taken before those checks: these counters are then kept by host_check() as the pairs are handed out */
	if (a->has_check)
		memset(tmp, 0, sizeof tmp);
	if (st->has_codes)
		rc = pb_assemble_host_codes(a->ctx, &cfg, st->n, st->f_data, st->f_off, st->r_data, st->r_off, st->res, st->nt, st->code, st->stride, cnt);
	else
		rc = pb_assemble_host(a->ctx, &cfg, st->n, st->f_data, st->f_off, st->r_data, st->r_off, st->res, st->nt, st->p, st->stride, cnt);
	if (rc != PB_OK)
		return false;
	if (a->has_check) {
		tmp[PB_C_OK] = 0;
		tmp[PB_C_LONGEST] = 0;
		memset(tmp + PB_C_OVERLAPS, 0, (PB_NCOUNTERS - PB_C_OVERLAPS) * sizeof(int64_t));
		pb_counters_merge(a->counters, tmp);
	}
	return true;
}

/* Expand device record i of a stage into a->result (pandaseq-common.h:277-330). */
static const panda_result_seq *publish(PandaAssembler a, const struct pb_stage *st, size_t i, const panda_seq_identifier *id,
                                       const panda_qual *fwd, size_t flen, const panda_qual *rev, size_t rlen) {
	const pb_pair_result *r = &st->res[i];
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
	const uint8_t *nt = st->nt + i * (st->stride / 2);	/* 4 bit per base, base 2k in the low nibble */
	const size_t len = r->seq_len < st->stride ? r->seq_len : st->stride;
	if (st->has_codes) {
		const uint16_t *code = st->code + i * st->stride;
		const double *table = a->ptable;
		panda_result *seq = a->result_seq;
		size_t k = 0;
		for (; k + 2 <= len; k += 2) {		/* two bases per packed byte */
			const unsigned b = nt[k >> 1], c0 = code[k], c1 = code[k + 1];
			seq[k].nt = (panda_nt) (b & 0x0F);
			seq[k].p = table[c0 < PB_POSTERIOR_CODES ? c0 : 0];
			seq[k + 1].nt = (panda_nt) (b >> 4);
			seq[k + 1].p = table[c1 < PB_POSTERIOR_CODES ? c1 : 0];
		}
		if (k < len) {
			seq[k].nt = (panda_nt) (nt[k >> 1] & 0x0F);
			seq[k].p = table[code[k] < PB_POSTERIOR_CODES ? code[k] : 0];
		}
	} else {
		const double *p = st->p + i * st->stride;
		for (size_t k = 0; k < len; k++) {
			a->result_seq[k].nt = (panda_nt) ((nt[k >> 1] >> ((k & 1) * 4)) & 0x0F);
			a->result_seq[k].p = p[k];
		}
	}
	return out;
}

const panda_result_seq *panda_assembler_assemble(PandaAssembler a, panda_seq_identifier *id,
                                                 const panda_qual *forward, size_t forward_length,
                                                 const panda_qual *reverse, size_t reverse_length) {
	struct pb_stage *st = &a->call;
	assert(forward_length <= PB_MAX_LEN);
	assert(reverse_length <= PB_MAX_LEN);
	if (!host_precheck(a, id, forward, forward_length, reverse, reverse_length))
		return NULL;
	stage_begin(st);
	if (!stage_push(st, id, forward, forward_length, reverse, reverse_length) || !run_batch(a, st))
		return NULL;
	if (st->res[0].status == PB_PAIR_NOALGN && a->noalgn != NULL)
		a->noalgn(a, id, forward, forward_length, reverse, reverse_length, a->noalgn_data);
	if (st->res[0].status != PB_PAIR_OK)
		return NULL;
	const panda_result_seq *out = publish(a, st, 0, id, forward, forward_length, reverse, reverse_length);
	return host_check(a, out) ? out : NULL;
}

size_t panda_assembler_assemble_batch(PandaAssembler a, size_t n, const panda_seq_identifier *ids,
                                      const panda_qual *const *forward, const size_t *forward_length,
                                      const panda_qual *const *reverse, const size_t *reverse_length,
                                      PandaOutputSeq output, void *output_data) {
	struct pb_stage *st = &a->call;
	size_t fb = 0, rb = 0, accepted = 0;
	for (size_t i = 0; i < n; i++) {
		if (forward_length[i] > PB_MAX_LEN || reverse_length[i] > PB_MAX_LEN) {
			pb_set_error("pair %zu longer than PANDA_MAX_LEN", i);
			return (size_t) -1;
		}
		fb += forward_length[i];
		rb += reverse_length[i];
	}
	stage_begin(st);
	if (!reserve(st, n, fb, rb, 0, 0, 0))
		return (size_t) -1;
	/* pairs a module's pre-check rejects are not staged; map[k] = caller's index of staged pair k */
	size_t *map = a->modules_length ? malloc((n ? n : 1) * sizeof(size_t)) : NULL;
	if (a->modules_length && map == NULL)
		return (size_t) -1;
	for (size_t i = 0; i < n; i++) {
		if (!host_precheck(a, ids ? &ids[i] : NULL, forward[i], forward_length[i], reverse[i], reverse_length[i]))
			continue;
		if (map)
			map[st->n] = i;
		if (!stage_push(st, ids ? &ids[i] : NULL, forward[i], forward_length[i], reverse[i], reverse_length[i])) {
			free(map);
			return (size_t) -1;
		}
	}
	if (!run_batch(a, st)) {
		free(map);
		return (size_t) -1;
	}
	for (size_t k = 0; k < st->n; k++) {
		const size_t i = map ? map[k] : k;
		const panda_seq_identifier *id = ids ? &ids[i] : NULL;
		if (st->res[k].status == PB_PAIR_NOALGN && a->noalgn != NULL)
			a->noalgn(a, id, forward[i], forward_length[i], reverse[i], reverse_length[i], a->noalgn_data);
		if (st->res[k].status != PB_PAIR_OK)
			continue;
		const panda_result_seq *out = publish(a, st, k, id, forward[i], forward_length[i], reverse[i], reverse_length[i]);
		if (!host_check(a, out))
			continue;
		accepted++;
		if (output != NULL)
			output(out, output_data);
	}
	free(map);
	return accepted;
}

/* the next accepted pair of a stage that has been through the device, or NULL when it is used up */
static const panda_result_seq *stage_next_result(PandaAssembler a, struct pb_stage *st) {
	while (st->pos < st->n) {
		const size_t i = st->pos++;
		const panda_qual *f = st->f_data + st->f_off[i], *r = st->r_data + st->r_off[i];
		const size_t fl = (size_t) (st->f_off[i + 1] - st->f_off[i]), rl = (size_t) (st->r_off[i + 1] - st->r_off[i]);
		if (st->res[i].status == PB_PAIR_NOALGN && a->noalgn != NULL)
			a->noalgn(a, &st->ids[i], f, fl, r, rl, a->noalgn_data);
		if (st->res[i].status == PB_PAIR_OK) {
			const panda_result_seq *out = publish(a, st, i, &st->ids[i], f, fl, r, rl);
			if (host_check(a, out))
				return out;
		}
	}
	return NULL;
}

/* Pull up to `limit` pairs from `src`'s source into `st` (pre-checks run here, against `a`'s modules and counters).
 * Returns false when staging failed; *dry is set when the source ended. */
static bool stage_fill(PandaAssembler a, PandaAssembler src, struct pb_stage *st, size_t limit, bool *dry) {
	stage_adopt(st);
	stage_begin(st);
	if (!reserve(st, limit, limit * 160, limit * 160, 0, 0, 0))
		return false;
	while (st->n < limit) {
		panda_seq_identifier id;
		const panda_qual *f, *r;
		size_t fl, rl;
		if (!src->next(&id, &f, &fl, &r, &rl, src->next_data)) {
			*dry = true;
			break;
		}
		assert(fl <= PB_MAX_LEN);
		assert(rl <= PB_MAX_LEN);
		if (!host_precheck(a, &id, f, fl, r, rl))
			continue;	/* counted and rejected on the host: not staged */
		if (!stage_push(st, &id, f, fl, r, rl))
			return false;
	}
	return true;
}

const panda_result_seq *panda_assembler_next(PandaAssembler a) {
	struct pb_stage *st = &a->stream;
	if (a->next == NULL)
		return NULL;
	for (;;) {
		/* hand out what the last launch produced, in input order */
		const panda_result_seq *out = stage_next_result(a, st);
		if (out != NULL)
			return out;
		if (a->source_dry || a->failed)
			return NULL;
		/* refill */
		if (!stage_fill(a, a, st, a->next_batch, &a->source_dry) || !run_batch(a, st)) {
			/* NULL is also the end of the stream, so the failure is made to stick: no later call pulls further pairs and
			 * leaves a hole; pb_last_error() has the reason */
			a->failed = true;
			st->n = st->pos = 0;
			return NULL;
		}
		if (st->n == 0 && a->source_dry)
			return NULL;
	}
}

/* ---- panda_run_pool (pool.c:110-181) ------------------------------------------------------------------------------------
 * The reference fans one assembler out over `threads` workers, each with a clone made by panda_assembler_copy_configuration,
 * all pulling pairs from one mutex-protected source, each handing its results to `output` from its own thread, in no
 * particular order (pandaseq.1:208).  The same here, with a GPU behind every worker: worker w gets device w mod D, D = the
 * visible GPUs (PANDASEQ_B200_DEVICES caps it), pulls a batch while it holds the source, runs it on its device and hands the
 * results out while the other workers pull and compute.  The workers' counters are merged into `assembler` at the end (the STAT
 * merge), so the caller's getters see the whole run.  The mux is an opaque handle this library
 * never creates, and is ignored.  Ownership follows the reference: the assembler is consumed (unref),
 * output_destroy(output_data) is called at the end.  Returns whether any pair was read. */
struct pool_shared {
	PandaAssembler source;
	pthread_mutex_t source_mutex;
	bool dry, stop, failed;
	PandaOutputSeq output;
	void *output_data;
};
struct pool_worker {
	struct pool_shared *shared;
	PandaAssembler self;
	pthread_t tid;
	int device;	/* -1: one GPU, the workers run wherever the scheduler puts them */
};

static void *pool_work(void *arg) {
	struct pool_worker *w = arg;
	struct pool_shared *sh = w->shared;
	PandaAssembler a = w->self;
	struct pb_stage *st = &a->stream;
	if (w->device >= 0)
		pb_bind_thread_near_device(w->device);	/* only with several GPUs: a worker stays on the socket of its own */
	for (;;) {
		bool ok, dry = false;
		pthread_mutex_lock(&sh->source_mutex);
		if (sh->dry || sh->stop || sh->failed) {
			pthread_mutex_unlock(&sh->source_mutex);
			break;
		}
		ok = stage_fill(a, sh->source, st, a->next_batch, &dry);
		if (dry)
			sh->dry = true;
		if (!ok)
			sh->failed = true;
		pthread_mutex_unlock(&sh->source_mutex);
		if (!ok || !run_batch(a, st)) {
			pthread_mutex_lock(&sh->source_mutex);
			sh->failed = true;
			pthread_mutex_unlock(&sh->source_mutex);
			break;
		}
		const panda_result_seq *result;
		while ((result = stage_next_result(a, st)) != NULL) {
			if (sh->output != NULL && !sh->output(result, sh->output_data)) {
				pthread_mutex_lock(&sh->source_mutex);
				sh->stop = true;
				pthread_mutex_unlock(&sh->source_mutex);
				break;
			}
		}
	}
	return NULL;
}

bool panda_run_pool(int threads, PandaAssembler assembler, PandaMux mux, PandaOutputSeq output, void *output_data, PandaDestroy output_destroy) {
	bool some_seqs;
	(void) mux;
	if (assembler == NULL) {
		if (output_destroy != NULL)
			output_destroy(output_data);
		return false;
	}
	if (threads > 64)
		threads = 64;
	if (threads <= 1 || assembler->next == NULL) {
		const panda_result_seq *result;
		while ((result = panda_assembler_next(assembler)) != NULL) {
			if (output != NULL && !output(result, output_data))
				break;
		}
	} else {
		struct pool_shared sh;
		struct pool_worker *workers = calloc((size_t) threads, sizeof *workers);
		int devices = pb_device_count(), started = 0;
		const char *env = getenv("PANDASEQ_B200_DEVICES");
		if (env != NULL && atoi(env) > 0 && atoi(env) < devices)
			devices = atoi(env);
		if (devices < 1)
			devices = 1;
		memset(&sh, 0, sizeof sh);
		sh.source = assembler;
		sh.output = output;
		sh.output_data = output_data;
		pthread_mutex_init(&sh.source_mutex, NULL);
		for (int k = 0; workers != NULL && k < threads; k++) {
			pb_context *ctx = NULL;
			PandaAssembler c;
			if (pb_device_context(k % devices, &ctx) != PB_OK)
				break;
			c = panda_assembler_new_kmer(NULL, NULL, NULL, assembler->logger, assembler->num_kmers);
			if (c == NULL)
				break;
			panda_assembler_copy_configuration(c, assembler);
			c->ctx = ctx;
			c->next_batch = assembler->next_batch > 32768 ? 32768 : assembler->next_batch;
			/* the worker's stage is set up here, before any worker runs: page-locked allocations made while other threads
			 * copy and launch stall them all (and a stage a previous run_pool gave up is simply taken over) */
			stage_adopt(&c->stream);
			if (reserve(&c->stream, c->next_batch, c->next_batch * 160, c->next_batch * 160, 0, 0, 0)) {
				grow_pinned((void **) &c->stream.res, &c->stream.cap_res, c->next_batch, sizeof(pb_pair_result), 0);
				grow_pinned((void **) &c->stream.nt, &c->stream.cap_nt, c->next_batch * 160, 1, 0);
				grow_pinned((void **) &c->stream.code, &c->stream.cap_code, c->next_batch * 320, sizeof(uint16_t), 0);
			}
			c->noalgn = assembler->noalgn;		/* called from the worker's thread, as the reference's clones do; not owned */
			c->noalgn_data = assembler->noalgn_data;
			c->noalgn_destroy = NULL;
			workers[k].shared = &sh;
			workers[k].self = c;
			workers[k].device = devices > 1 ? k % devices : -1;
			if (pthread_create(&workers[k].tid, NULL, pool_work, &workers[k]) != 0) {
				panda_assembler_unref(c);
				break;
			}
			started++;
		}
		if (started == 0) {		/* no worker could be set up: the plain loop still works */
			const panda_result_seq *result;
			while ((result = panda_assembler_next(assembler)) != NULL) {
				if (output != NULL && !output(result, output_data))
					break;
			}
		}
		for (int k = 0; k < started; k++) {
			PandaAssembler c = workers[k].self;
			pthread_join(workers[k].tid, NULL);
			pb_counters_merge(assembler->counters, c->counters);
			for (size_t it = 0; it < c->modules_length && it < assembler->modules_length; it++)
				assembler->rejected[it] += c->rejected[it];
			c->noalgn = NULL;
			panda_assembler_unref(c);
		}
		if (sh.failed)
			assembler->failed = true;
		pthread_mutex_destroy(&sh.source_mutex);
		free(workers);
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
