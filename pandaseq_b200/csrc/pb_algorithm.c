/* pb_algorithm.c -- PandaAlgorithm objects for the device-scored algorithms (the reference's full registry of seven).
 *
 * Keeps the reference's algorithm plug-in surface (pandaseq-algorithm.h:33-228,
 * algo.c:27-133): refcounted instances of a class, private data behind
 * panda_algorithm_data(), a sorted registry, per-class parameter accessors.
 * What differs: the class's overlap_probability pointer is not a CPU scorer.
 * Scoring candidate overlaps is the device kernel's job (pb_kernels.cuh);
 * the host pointer exists only so the struct keeps its shape and reports NaN
 * if someone calls it.  match_probability IS evaluated on the host, because
 * it is how the 2x48x48 reconstruction table is tabulated before a launch
 * (pb_luts.c) and what panda_algorithm_quality_compare returns.
 */
#define _GNU_SOURCE
#include "pb_internal.h"
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct sb_private {
	double q;
};
struct pear_private {
	double random_base;
};

/* Device-only: see the header comment. */
static double overlap_on_device_only(void *data, const panda_qual *f, size_t fl, const panda_qual *r, size_t rl, size_t overlap) {
	(void) data; (void) f; (void) fl; (void) r; (void) rl; (void) overlap;
	return NAN;
}

static double sb_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_SIMPLE_BAYES, match, a, b);
}
static double pear_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_PEAR, match, a, b);
}
static double rdp_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_RDP_MLE, match, a, b);
}
static double flash_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_FLASH, match, a, b);
}
static double ea_util_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_EA_UTIL, match, a, b);
}
static double stitch_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_STITCH, match, a, b);
}
static double uparse_match(void *data, bool match, char a, char b) {
	(void) data;
	return pb_host_match_probability(PB_UPARSE, match, a, b);
}

/* "-A name:argument" parsing, same acceptance rules as the reference's from_string functions. */
static bool parse_probability(const char *text, const char *what, double *out) {
	char *end;
	errno = 0;
	*out = strtod(text, &end);
	if (errno == ERANGE || *end != '\0') {
		fprintf(stderr, "Cannot parse value: %s\n", text);
		return false;
	}
	if (*out < 0 || *out > 1) {
		fprintf(stderr, "%s %f is not a probability.\n", what, *out);
		return false;
	}
	return true;
}

static PandaAlgorithm sb_create(const char *arg) {	/* algo_simple_bayes.c:77-98 */
	double q;
	PandaAlgorithm algo;
	if (arg == NULL)
		return panda_algorithm_simple_bayes_new();
	if (!parse_probability(arg, "Error estimation", &q))
		return NULL;
	algo = panda_algorithm_simple_bayes_new();
	panda_algorithm_simple_bayes_set_error_estimation(algo, q);
	return algo;
}

static PandaAlgorithm pear_create(const char *arg) {	/* algo_pear.c:70-91 */
	double p;
	PandaAlgorithm algo;
	if (arg == NULL)
		return panda_algorithm_pear_new();
	if (!parse_probability(arg, "Random base", &p))
		return NULL;
	algo = panda_algorithm_pear_new();
	panda_algorithm_pear_set_random_base_log_p(algo, log(p));
	return algo;
}

static PandaAlgorithm rdp_create(const char *arg) {	/* algo_rdp_mle.c:76-82 */
	if (arg == NULL || arg[0] == '\0')
		return panda_algorithm_rdp_mle_new();
	return NULL;
}

static PandaAlgorithm flash_create(const char *arg) {	/* algo_flash.c:82-89 */
	if (arg != NULL) {
		fprintf(stderr, "No arguments allowed: %s\n", arg);
		return NULL;
	}
	return panda_algorithm_flash_new();
}

static PandaAlgorithm ea_util_create(const char *arg) {	/* algo_ea_util.c:69-77 */
	if (arg != NULL) {
		fprintf(stderr, "No arguments allowed: %s\n", arg);
		return NULL;
	}
	return panda_algorithm_ea_util_new();
}

static PandaAlgorithm stitch_create(const char *arg) {	/* algo_stitch.c:68-77 */
	if (arg != NULL) {
		fprintf(stderr, "No arguments allowed: %s\n", arg);
		return NULL;
	}
	return panda_algorithm_stitch_new();
}

static PandaAlgorithm uparse_create(const char *arg) {	/* algo_uparse.c:77-98 */
	double q;
	PandaAlgorithm algo;
	if (arg == NULL)
		return panda_algorithm_uparse_new();
	if (!parse_probability(arg, "Error estimation", &q))
		return NULL;
	algo = panda_algorithm_uparse_new();
	panda_algorithm_uparse_set_error_estimation(algo, q);
	return algo;
}

/* qual_nn_simple_bayesian is the literal -1.38629 in every class (table.h via tablebuilder.c:124). */
#define PB_QUAL_NN (-1.38629)

const struct panda_algorithm_class panda_algorithm_simple_bayes_class = {
	sizeof(struct sb_private), "simple_bayesian", sb_create, NULL, overlap_on_device_only, sb_match, PB_QUAL_NN
};
const struct panda_algorithm_class panda_algorithm_pear_class = {
	sizeof(struct pear_private), "pear", pear_create, NULL, overlap_on_device_only, pear_match, PB_QUAL_NN
};
const struct panda_algorithm_class panda_algorithm_rdp_mle_class = {
	0, "rdp_mle", rdp_create, NULL, overlap_on_device_only, rdp_match, PB_QUAL_NN
};
const struct panda_algorithm_class panda_algorithm_flash_class = {
	0, "flash", flash_create, NULL, overlap_on_device_only, flash_match, PB_QUAL_NN
};

const struct panda_algorithm_class panda_algorithm_ea_util_class = {
	0, "ea_util", ea_util_create, NULL, overlap_on_device_only, ea_util_match, PB_QUAL_NN
};
const struct panda_algorithm_class panda_algorithm_stitch_class = {
	0, "stitch", stitch_create, NULL, overlap_on_device_only, stitch_match, PB_QUAL_NN
};
const struct panda_algorithm_class panda_algorithm_uparse_class = {
	sizeof(struct sb_private), "uparse", uparse_create, NULL, overlap_on_device_only, uparse_match, PB_QUAL_NN
};

/* ---- instances ---------------------------------------------------------- */

PandaAlgorithm panda_algorithm_new(PandaAlgorithmClass clazz) {
	PandaAlgorithm a = calloc(1, sizeof(struct panda_algorithm) + clazz->data_size);
	if (a == NULL)
		return NULL;
	pthread_mutex_init(&a->mutex, NULL);
	a->refcnt = 1;
	a->clazz = clazz;
	return a;
}

PandaAlgorithmClass panda_algorithm_class(PandaAlgorithm algo) {
	return algo->clazz;
}

void *panda_algorithm_data(PandaAlgorithm algo) {
	return &algo->end;
}

bool panda_algorithm_is_a(PandaAlgorithm algo, PandaAlgorithmClass clazz) {
	return algo != NULL && algo->clazz == clazz;
}

double panda_algorithm_quality_compare(PandaAlgorithm algorithm, const panda_qual *a, const panda_qual *b) {
	return algorithm->clazz->match_probability(panda_algorithm_data(algorithm), (a->nt & b->nt) != '\0', a->qual, b->qual);
}

PandaAlgorithm panda_algorithm_ref(PandaAlgorithm algo) {
	pthread_mutex_lock(&algo->mutex);
	algo->refcnt++;
	pthread_mutex_unlock(&algo->mutex);
	return algo;
}

void panda_algorithm_unref(PandaAlgorithm algo) {
	size_t left;
	if (algo == NULL)
		return;
	pthread_mutex_lock(&algo->mutex);
	left = --algo->refcnt;
	pthread_mutex_unlock(&algo->mutex);
	if (left != 0)
		return;
	pthread_mutex_destroy(&algo->mutex);
	if (algo->clazz->data_destroy != NULL)
		algo->clazz->data_destroy(panda_algorithm_data(algo));
	free(algo);
}

/* ---- per-class constructors and parameters ------------------------------- */

PandaAlgorithm panda_algorithm_simple_bayes_new(void) {
	PandaAlgorithm a = panda_algorithm_new(&panda_algorithm_simple_bayes_class);
	panda_algorithm_simple_bayes_set_error_estimation(a, 0.36);
	return a;
}

double panda_algorithm_simple_bayes_get_error_estimation(PandaAlgorithm algorithm) {
	if (!panda_algorithm_is_a(algorithm, &panda_algorithm_simple_bayes_class))
		return -1;
	return ((struct sb_private *) panda_algorithm_data(algorithm))->q;
}

void panda_algorithm_simple_bayes_set_error_estimation(PandaAlgorithm algorithm, double q) {
	if (q > 0 && q < 1 && panda_algorithm_is_a(algorithm, &panda_algorithm_simple_bayes_class))
		((struct sb_private *) panda_algorithm_data(algorithm))->q = q;
}

PandaAlgorithm panda_algorithm_pear_new(void) {
	PandaAlgorithm a = panda_algorithm_new(&panda_algorithm_pear_class);
	panda_algorithm_pear_set_random_base_log_p(a, log(0.25));
	return a;
}

double panda_algorithm_pear_get_random_base_log_p(PandaAlgorithm algorithm) {
	if (!panda_algorithm_is_a(algorithm, &panda_algorithm_pear_class))
		return 1;
	return ((struct pear_private *) panda_algorithm_data(algorithm))->random_base;
}

void panda_algorithm_pear_set_random_base_log_p(PandaAlgorithm algorithm, double log_p) {
	if (panda_algorithm_is_a(algorithm, &panda_algorithm_pear_class))
		((struct pear_private *) panda_algorithm_data(algorithm))->random_base = log_p;
}

PandaAlgorithm panda_algorithm_rdp_mle_new(void) {
	return panda_algorithm_new(&panda_algorithm_rdp_mle_class);
}

PandaAlgorithm panda_algorithm_flash_new(void) {
	return panda_algorithm_new(&panda_algorithm_flash_class);
}

PandaAlgorithm panda_algorithm_ea_util_new(void) {
	return panda_algorithm_new(&panda_algorithm_ea_util_class);
}

PandaAlgorithm panda_algorithm_stitch_new(void) {
	return panda_algorithm_new(&panda_algorithm_stitch_class);
}

PandaAlgorithm panda_algorithm_uparse_new(void) {
	PandaAlgorithm a = panda_algorithm_new(&panda_algorithm_uparse_class);
	panda_algorithm_uparse_set_error_estimation(a, 0.36);
	return a;
}

double panda_algorithm_uparse_get_error_estimation(PandaAlgorithm algorithm) {
	if (!panda_algorithm_is_a(algorithm, &panda_algorithm_uparse_class))
		return -1;
	return ((struct sb_private *) panda_algorithm_data(algorithm))->q;
}

void panda_algorithm_uparse_set_error_estimation(PandaAlgorithm algorithm, double q) {
	if (q > 0 && q < 1 && panda_algorithm_is_a(algorithm, &panda_algorithm_uparse_class))
		((struct sb_private *) panda_algorithm_data(algorithm))->q = q;
}

/* Class identity -> device scorer id + private data.  Unknown classes are refused:
 * a host function pointer cannot run in the kernel and there is no CPU path. */
int pb_algorithm_fill_config(PandaAlgorithm algo, pb_config *cfg) {
	if (panda_algorithm_is_a(algo, &panda_algorithm_simple_bayes_class)) {
		cfg->algo = PB_SIMPLE_BAYES;
		cfg->sb_q = ((struct sb_private *) panda_algorithm_data(algo))->q;
	} else if (panda_algorithm_is_a(algo, &panda_algorithm_pear_class)) {
		cfg->algo = PB_PEAR;
		cfg->pear_random_base = ((struct pear_private *) panda_algorithm_data(algo))->random_base;
	} else if (panda_algorithm_is_a(algo, &panda_algorithm_rdp_mle_class)) {
		cfg->algo = PB_RDP_MLE;
	} else if (panda_algorithm_is_a(algo, &panda_algorithm_flash_class)) {
		cfg->algo = PB_FLASH;
	} else if (panda_algorithm_is_a(algo, &panda_algorithm_ea_util_class)) {
		cfg->algo = PB_EA_UTIL;
	} else if (panda_algorithm_is_a(algo, &panda_algorithm_stitch_class)) {
		cfg->algo = PB_STITCH;
	} else if (panda_algorithm_is_a(algo, &panda_algorithm_uparse_class)) {
		cfg->algo = PB_UPARSE;
		cfg->sb_q = ((struct sb_private *) panda_algorithm_data(algo))->q;
	} else {
		pb_set_error("algorithm class '%s' has no device scorer", (algo && algo->clazz && algo->clazz->name) ? algo->clazz->name : "?");
		return -1;
	}
	return 0;
}

/* ---- registry (algo.c:85-133) --------------------------------------------- */

PandaAlgorithmClass *panda_algorithms = NULL;
size_t panda_algorithms_length = 0;
static size_t registry_capacity = 0;
static pthread_mutex_t registry_lock = PTHREAD_MUTEX_INITIALIZER;

static int by_name(const void *x, const void *y) {
	PandaAlgorithmClass const *a = x, *b = y;
	return strcmp((*a)->name, (*b)->name);
}

void panda_algorithm_register(PandaAlgorithmClass clazz) {
	pthread_mutex_lock(&registry_lock);
	for (size_t i = 0; i < panda_algorithms_length; i++) {
		if (panda_algorithms[i] == clazz) {
			pthread_mutex_unlock(&registry_lock);
			return;
		}
	}
	if (panda_algorithms_length == registry_capacity) {
		size_t grown = registry_capacity ? registry_capacity * 2 : 8;
		PandaAlgorithmClass *bigger = realloc(panda_algorithms, grown * sizeof *bigger);
		if (bigger == NULL) {
			pthread_mutex_unlock(&registry_lock);
			return;
		}
		panda_algorithms = bigger;
		registry_capacity = grown;
	}
	panda_algorithms[panda_algorithms_length++] = clazz;
	qsort(panda_algorithms, panda_algorithms_length, sizeof *panda_algorithms, by_name);
	pthread_mutex_unlock(&registry_lock);
}

__attribute__((constructor))
static void register_builtin_algorithms(void) {
	panda_algorithm_register(&panda_algorithm_ea_util_class);
	panda_algorithm_register(&panda_algorithm_flash_class);
	panda_algorithm_register(&panda_algorithm_pear_class);
	panda_algorithm_register(&panda_algorithm_rdp_mle_class);
	panda_algorithm_register(&panda_algorithm_simple_bayes_class);
	panda_algorithm_register(&panda_algorithm_stitch_class);
	panda_algorithm_register(&panda_algorithm_uparse_class);
}
