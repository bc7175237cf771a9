/* pb_luts.c -- host-side construction of the log-probability LUTs and of the
 * parameter block the kernels read.
 *
 * The reference generates table.c at build time with mktable (mktable.c:23-155)
 * and prints every entry with "%g" (tablebuilder.c:86,124,147), so the constants
 * the CPU path uses are 6-significant-digit decimals, not the exact formulas.
 * The same formulas and the same decimal round trip are applied here at library
 * load, once; tests/test_tables.py checks the result entry-by-entry against the
 * reference's generated table.c.
 */
#define _GNU_SOURCE
#include "pb_internal.h"
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char pb_error_buf[512];

void pb_set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(pb_error_buf, sizeof pb_error_buf, fmt, ap);
	va_end(ap);
}

const char *pb_last_error(void) {
	return pb_error_buf;
}

/* "%g" then strtod: what printing a constant into table.c and compiling it does. */
static double six_digits(double v) {
	char text[48];
	snprintf(text, sizeof text, "%g", v);
	return strtod(text, NULL);
}

/* prob.h:21 */
static double err_prob(int phred) {
	return pow(10.0, (-(double) phred) / 10.0);
}

typedef double (*pair_formula) (double p, double q);

/* mktable.c:23-31 */
static double sb_same(double p, double q) {
	return (1 - p) * (1 - q) + p * q / 3;
}
/* mktable.c:33-41 (and mismatch_rdp, mktable.c:84-92, which is the same expression) */
static double sb_diff(double p, double q) {
	return (1 - p) * q / 3 + (1 - q) * p / 3 + 2 * p * q / 9;
}
/* mktable.c:43-51 */
static double pear_same(double p, double q) {
	return (1 - (1 - q) * p / 3 - (1 - p) * q / 3 - 2 * (1 - p) * (1 - q) / 9);
}
/* mktable.c:53-61 */
static double pear_diff(double p, double q) {
	return (1 - p) * q / 3 + (1 - q) * p / 3 + p * q / 2;
}
/* mktable.c:94-104 */
static double rdp_assembled_diff(double p, double q) {
	double smaller = (p <= q) ? p : q;
	double v = 1 - (smaller - p * q / 3.0) / (p + q - 4.0 / 3.0 * p * q);
	return (v == 0) ? DBL_MIN : v;
}

/* mktable.c:106-117 */
static double uparse_same(double p, double q) {
	double v = 1 - p * q / (1 - p - q + 4 * p * q / 3);
	return (v <= 0) ? DBL_MIN : v;
}
/* mktable.c:119-131 */
static double uparse_diff(double p, double q) {
	double v = 1 - (p + q / 3) / (p + q - 4 * p * q / 3);
	return (v <= 0) ? DBL_MIN : v;
}

static void tabulate(double dst[PB_NQ][PB_NQ], pair_formula f) {
	for (int a = 0; a < PB_NQ; a++) {
		double pa = err_prob(a);
		for (int b = 0; b < PB_NQ; b++)
			dst[a][b] = six_digits(log(f(pa, err_prob(b))));	/* tablebuilder.c:154-166, log_output */
	}
}

static pb_tables the_tables;
static pthread_once_t the_tables_once = PTHREAD_ONCE_INIT;

static void make_tables(void) {
	pb_tables *t = &the_tables;
	t->qual_nn = six_digits(log(0.25));	/* mktable.c:141; header constant printed by tablebuilder.c:124 */
	tabulate(t->match_sb, sb_same);
	tabulate(t->mismatch_sb, sb_diff);
	tabulate(t->match_pear, pear_same);
	tabulate(t->mismatch_pear, pear_diff);
	tabulate(t->mismatch_rdp, sb_diff);
	tabulate(t->mismatch_rdp_asm, rdp_assembled_diff);
	tabulate(t->match_uparse, uparse_same);
	tabulate(t->mismatch_uparse, uparse_diff);
	for (int k = 0; k < PB_NQ; k++) {
		double p = err_prob(k);
		t->score[k] = six_digits(p == 1 ? -2.0 : log(1.0 - p));	/* mktable.c:63-73 */
		t->score_err[k] = six_digits(log(p));	/* mktable.c:75-82 */
	}
}

const pb_tables *pb_get_tables(void) {
	pthread_once(&the_tables_once, make_tables);
	return &the_tables;
}

/* prob.h:23 */
static int clamp_phred(char q) {
	return q > PB_PHREDMAX ? PB_PHREDMAX : (q < 0 ? 0 : q);
}

/* The per-base posterior each algorithm assigns during reconstruction:
 * algo_simple_bayes.c:68-75, algo_pear.c:61-68, algo_rdp_mle.c:29-41, algo_flash.c:62-80, and the three below.
 * Every one of them is a function of (match, clamp(a), clamp(b)) only, which is what
 * lets the device use one 2x48x48 table for all of them. */
double pb_host_match_probability(int algo, bool match, char a, char b) {
	const pb_tables *t = pb_get_tables();
	int qa = clamp_phred(a), qb = clamp_phred(b);
	switch (algo) {
	case PB_SIMPLE_BAYES:
		return match ? t->match_sb[qa][qb] : t->mismatch_sb[qa][qb];
	case PB_PEAR:
		return match ? t->match_pear[qa][qb] : t->mismatch_pear[qa][qb];
	case PB_RDP_MLE:
		if (match)
			return t->score[(a >= b) ? qa : qb];
		return t->mismatch_rdp_asm[qa][qb];
	case PB_EA_UTIL:	/* algo_ea_util.c:58-67 */
		return t->score[(a > b) ? qa : qb];
	case PB_STITCH:		/* algo_stitch.c:58-66 */
		return match ? t->match_sb[qa][qb] : t->mismatch_sb[qa][qb];
	case PB_UPARSE:		/* algo_uparse.c:68-75 */
		return match ? t->match_uparse[qa][qb] : t->mismatch_uparse[qa][qb];
	case PB_FLASH:
		if (match)
			return t->score[(a > b) ? qa : qb];
		else {
			int d = qa - qb;
			if (d < 0)
				d = -d;
			return t->score[d < 2 ? 2 : d];
		}
	}
	return NAN;
}

void pb_config_default(pb_config *cfg, int algo) {
	memset(cfg, 0, sizeof *cfg);
	cfg->algo = algo;
	cfg->minoverlap = 2;		/* assembler_support.c:96 */
	cfg->maxoverlap = 0;		/* assembler_support.c:95 */
	cfg->num_kmers = PANDA_DEFAULT_NUM_KMERS;
	cfg->threshold = log(0.6);	/* assembler_support.c:76 */
	cfg->primer_penalty = 0;	/* assembler_support.c:97 */
	cfg->sb_q = 0.36;		/* algo_simple_bayes.c:113 */
	cfg->pear_random_base = log(0.25);	/* algo_pear.c:106 */
}

void pb_counters_merge(int64_t *dst, const int64_t *src) {
	for (int i = 0; i < PB_NCOUNTERS; i++) {
		if (i == PB_C_LONGEST) {
			if (dst[i] < src[i])
				dst[i] = src[i];
		} else {
			dst[i] += src[i];
		}
	}
}

pb_status pb_build_device_params(const pb_config *cfg, pb_device_params *out) {
	const pb_tables *t = pb_get_tables();
	if (cfg->algo < PB_SIMPLE_BAYES || cfg->algo > PB_UPARSE) {
		pb_set_error("algorithm %d has no device scorer", cfg->algo);
		return PB_ERR_UNSUPPORTED;
	}
	if (cfg->num_kmers != 2) {
		pb_set_error("num_kmers=%ld: the reference's k-mer table indexing (assembler.c:94 vs :99) is only self-consistent for 2", (long) cfg->num_kmers);
		return PB_ERR_UNSUPPORTED;
	}
	if (cfg->minoverlap < 2 || cfg->minoverlap >= 2 * PB_MAX_LEN || cfg->maxoverlap < 0 || cfg->maxoverlap >= 2 * PB_MAX_LEN
	    || cfg->forward_primer_length < 0 || cfg->forward_primer_length >= PB_MAX_LEN
	    || cfg->reverse_primer_length < 0 || cfg->reverse_primer_length >= PB_MAX_LEN
	    || cfg->forward_trim < 0 || cfg->reverse_trim < 0 || cfg->forward_trim > 65535 || cfg->reverse_trim > 65535) {
		pb_set_error("configuration outside the ranges the reference's setters allow");
		return PB_ERR_ARGUMENT;
	}
	if (cfg->hang_forward_length < 0 || cfg->hang_forward_length >= PB_MAX_LEN || cfg->hang_reverse_length < 0
	    || cfg->hang_reverse_length >= PB_MAX_LEN || cfg->nfilters < 0 || cfg->nfilters > PB_MAX_FILTERS) {
		pb_set_error("overhang sequences must be shorter than %d, at most %d filters", PB_MAX_LEN, PB_MAX_FILTERS);
		return PB_ERR_ARGUMENT;
	}
	for (int k = 0; k < cfg->nfilters; k++) {
		const struct pb_filter *f = &cfg->filters[k];
		if (f->kind == PB_FILTER_PEAR_TEST) {       /* plugin_pear_test.c:97-100: the cut-off is a p-value */
			if (!(f->dvalue3 >= 0 && f->dvalue3 <= 1)) {
				pb_set_error("filter %d: pear_test cutoff out of range", k);
				return PB_ERR_ARGUMENT;
			}
			continue;
		}
		if (f->kind < PB_FILTER_NO_N || f->kind > PB_FILTER_MIN_PHRED || (f->kind == PB_FILTER_MIN_OVERLAPBITS && !(f->dvalue >= 0))
		    || (f->kind != PB_FILTER_MIN_OVERLAPBITS && f->ivalue < 0)) {
			pb_set_error("filter %d: unknown kind or value out of range", k);
			return PB_ERR_ARGUMENT;
		}
	}
	memset(out, 0, sizeof *out);
	out->hang_forward_length = (int32_t) cfg->hang_forward_length;
	out->hang_reverse_length = (int32_t) cfg->hang_reverse_length;
	out->hang_skip = cfg->hang_skip ? 1 : 0;
	out->hang_threshold = cfg->hang_threshold;
	for (int k = 0; k < cfg->hang_forward_length; k++)      /* hang.c:103-104 */
		out->hang_forward[cfg->hang_forward_length - k - 1] = (uint8_t) cfg->hang_forward[k] & 15;
	for (int k = 0; k < cfg->hang_reverse_length; k++)      /* hang.c:105-106 */
		out->hang_reverse[cfg->hang_reverse_length - k - 1] = (uint8_t) cfg->hang_reverse[k] & 15;
	out->nfilters = cfg->nfilters;
	for (int k = 0; k < cfg->nfilters; k++) {
		out->filters[k] = cfg->filters[k];
		if (cfg->filters[k].kind == PB_FILTER_MIN_PHRED)
			out->need_stage = 1;
	}
	out->algo = cfg->algo;
	out->minoverlap = (int32_t) cfg->minoverlap;
	out->maxoverlap = (int32_t) cfg->maxoverlap;
	out->forward_trim = (int32_t) cfg->forward_trim;
	out->reverse_trim = (int32_t) cfg->reverse_trim;
	out->forward_primer_length = (int32_t) cfg->forward_primer_length;
	out->reverse_primer_length = (int32_t) cfg->reverse_primer_length;
	out->post_primers = cfg->post_primers ? 1 : 0;
	out->threshold = cfg->threshold;
	out->primer_penalty = cfg->primer_penalty;
	out->qual_nn = t->qual_nn;
	out->pear_random_base = cfg->pear_random_base;
	{
		double q = cfg->sb_q;
		if (cfg->algo == PB_UPARSE) {	/* algo_uparse.c:126-135 */
			out->sb_pmatch = log(1 - q * q * (1 - 2 * q + 4 * q * q / 3));
			out->sb_pmismatch = log(1 - 4 * q / 3 / (2 * q - 4 * q * q / 3));
		} else {			/* algo_simple_bayes.c:126-135 */
			out->sb_pmatch = log(0.25 * (1 - 2 * q + q * q));
			out->sb_pmismatch = log((3 * q - 2 * q * q) / 18.0);
		}
	}
	for (int k = 0; k < PB_NQ; k++) {
		out->score[k] = t->score[k];
		out->score_err[k] = t->score_err[k];
	}
	for (int m = 0; m < 2; m++) {
		for (int a = 0; a < PB_NQM; a++) {
			for (int b = 0; b < PB_NQM; b++) {
				double v;
				if (a == PB_NQ && b == PB_NQ)
					v = t->qual_nn;	/* both reads masked: assembler.c:202-203 */
				else if (a == PB_NQ)
					v = t->score[b];	/* assembler.c:204-205, 235 */
				else if (b == PB_NQ)
					v = t->score[a];	/* assembler.c:206-207, 165 */
				else
					v = pb_host_match_probability(cfg->algo, m != 0, (char) a, (char) b);
				out->recon[m][a][b] = v;
			}
		}
	}
	for (int a = 0; a < PB_NQ; a++) {
		for (int b = 0; b < PB_NQ; b++) {
			switch (cfg->algo) {
			case PB_PEAR:	/* algo_pear.c:52,54 */
				out->over[1][a][b] = t->match_pear[a][b];
				out->over[0][a][b] = t->mismatch_pear[a][b];
				break;
			case PB_RDP_MLE:	/* algo_rdp_mle.c:68,70: the term is (table - qual_nn), formed before the add */
				out->over[1][a][b] = t->match_sb[a][b] - t->qual_nn;
				out->over[0][a][b] = t->mismatch_rdp[a][b] - t->qual_nn;
				break;
			default:
				break;
			}
		}
	}
	for (int i = 0; i < cfg->forward_primer_length; i++)
		out->forward_primer[i] = (uint8_t) cfg->forward_primer[i] & 0x0F;
	for (int i = 0; i < cfg->reverse_primer_length; i++)
		out->reverse_primer[i] = (uint8_t) cfg->reverse_primer[i] & 0x0F;
	return PB_OK;
}

/* plugin_pear_test.c:31-35: for every i the partial sums of its inner loop, in its order and with the same libm calls, so
 * that the device's product over i multiplies exactly the doubles the plugin's would.  Independent of alpha / beta / cutoff. */
void pb_build_pear_cdf(double *out) {
	for (size_t i = 0; i < PB_PEAR_ROWS; i++) {
		double sum = 0;
		double *row = out + i * PB_PEAR_COLS;
		row[0] = 0;
		for (size_t k = 0; k <= i; k++) {
			double i_choose_k = lgamma(i + 1) - lgamma(k + 1) - lgamma(i - k + 1);
			sum += exp(i_choose_k + k * log(0.25) + (i - k) * log(0.75));
			row[k + 1] = sum;
		}
		for (size_t l = i + 2; l < PB_PEAR_COLS; l++)
			row[l] = sum;
	}
}

size_t panda_max_len(void) {
	return PB_MAX_LEN;
}

size_t pb_layout_host(size_t n, const uint64_t *f_off, const uint64_t *r_off, uint32_t *rec_off16) {
	size_t total16 = 0;
	for (size_t i = 0; i < n; i++) {
		rec_off16[i] = (uint32_t) total16;
		total16 += pb_record_bytes((size_t) (f_off[i + 1] - f_off[i]), (size_t) (r_off[i + 1] - r_off[i])) / 16;
	}
	return total16 * 16;
}

pb_status pb_posterior_table(const pb_config *cfg, double *table) {
	pb_device_params *prm;
	pb_status st;
	if (cfg == NULL || table == NULL) {
		pb_set_error("pb_posterior_table: bad argument");
		return PB_ERR_ARGUMENT;
	}
	prm = malloc(sizeof *prm);
	if (prm == NULL)
		return PB_ERR_NOMEM;
	st = pb_build_device_params(cfg, prm);
	if (st == PB_OK)
		memcpy(table, prm->recon, sizeof prm->recon);
	free(prm);
	return st;
}

/* The packed layout of include/pandaseq_b200.h written on the host: what pb::pack_kernel does on the device. */
void pb_pack_host(size_t n, const panda_qual *f_data, const uint64_t *f_off, const panda_qual *r_data, const uint64_t *r_off,
                  uint8_t *reads, pb_pair_meta *meta) {
	size_t total16 = 0;
	for (size_t i = 0; i < n; i++) {
		const size_t F = (size_t) (f_off[i + 1] - f_off[i]), R = (size_t) (r_off[i + 1] - r_off[i]);
		const panda_qual *f = f_data + f_off[i], *r = r_data + r_off[i];
		const size_t fw = ((F + 7) / 8) * 4, rw = ((R + 7) / 8) * 4, fq = ((F + 3) / 4) * 4, bytes = pb_record_bytes(F, R);
		uint8_t *rec = reads + total16 * 16;
		memset(rec, 0, bytes);
		for (size_t k = 0; k < F; k++) {
			rec[k >> 1] |= (uint8_t) ((f[k].nt & 15) << ((k & 1) * 4));
			rec[fw + rw + k] = (uint8_t) f[k].qual;
		}
		for (size_t k = 0; k < R; k++) {      /* template order */
			const panda_qual *e = &r[R - 1 - k];
			rec[fw + (k >> 1)] |= (uint8_t) ((e->nt & 15) << ((k & 1) * 4));
			rec[fw + rw + fq + k] = (uint8_t) e->qual;
		}
		meta[i].off16 = (uint32_t) total16;
		meta[i].flen = (uint16_t) F;
		meta[i].rlen = (uint16_t) R;
		total16 += bytes / 16;
	}
}
