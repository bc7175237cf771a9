/* panda_oracle.c -- CPU restatement of the PANDAseq pair-assembly hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see panda_oracle.h).  Plain scalar C99, written
 * from the reference's behaviour as catalogued in SURVEY.md §8a; each function
 * names the reference file:line it follows.  It is deliberately a different
 * program shape from the reference (flat batch in, structure-of-arrays out, one
 * config struct instead of objects and function pointers) while computing the
 * same doubles in the same order, so that it can be compared bit-for-bit with
 * oracle/_ref/libpandaseq_ref.so and then serve as the checker for the CUDA path.
 */
#define _GNU_SOURCE
#include "panda_oracle.h"
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ tables */

/* tablebuilder.c:86,124,147 print every constant with "%g"; the compiler then
 * parses the 6-significant-digit decimal.  Same round trip here. */
static double g_round(double x) {
	char buf[64];
	snprintf(buf, sizeof buf, "%g", x);
	return strtod(buf, NULL);
}

/* prob.h:21 */
static double phred_p(int score) {
	return pow(10.0, (-(double) score) / 10.0);
}

/* mktable.c:23-104 -- the six matrix formulas and two array formulas. */
static double f_match_sb(double p, double q) { return (1 - p) * (1 - q) + p * q / 3; }
static double f_mismatch_sb(double p, double q) { return (1 - p) * q / 3 + (1 - q) * p / 3 + 2 * p * q / 9; }
static double f_match_pear(double p, double q) { return (1 - (1 - q) * p / 3 - (1 - p) * q / 3 - 2 * (1 - p) * (1 - q) / 9); }
static double f_mismatch_pear(double p, double q) { return (1 - p) * q / 3 + (1 - q) * p / 3 + p * q / 2; }
static double f_mismatch_rdp(double p, double q) { return ((1 - p) * q / 3 + (1 - q) * p / 3 + 2 * p * q / 9); }
static double f_mismatch_rdp_asm(double p, double q) {
	double lo = (p <= q) ? p : q;
	double v = 1 - (lo - p * q / 3.0) / (p + q - 4.0 / 3.0 * p * q);
	return (v == 0) ? DBL_MIN : v;
}

/* mktable.c:106-131 */
static double f_match_uparse(double p, double q) {
	double v = 1 - p * q / (1 - p - q + 4 * p * q / 3);
	return (v <= 0) ? DBL_MIN : v;
}
static double f_mismatch_uparse(double p, double q) {
	double v = 1 - (p + q / 3) / (p + q - 4 * p * q / 3);
	return (v <= 0) ? DBL_MIN : v;
}

static po_tables g_tables;
static pthread_once_t g_tables_once = PTHREAD_ONCE_INIT;

static void fill_matrix(double m[PO_NQ][PO_NQ], double (*f) (double, double)) {
	/* tablebuilder.c:154-183: entry = log(formula(P(x), P(y))), printed "%g". */
	for (int x = 0; x < PO_NQ; x++)
		for (int y = 0; y < PO_NQ; y++)
			m[x][y] = g_round(log(f(phred_p(x), phred_p(y))));
}

static void build_tables(void) {
	po_tables *t = &g_tables;
	t->qual_nn = g_round(log(0.25));	/* mktable.c:141 + tablebuilder.c:124 -> -1.38629 */
	fill_matrix(t->match_sb, f_match_sb);
	fill_matrix(t->mismatch_sb, f_mismatch_sb);
	fill_matrix(t->match_pear, f_match_pear);
	fill_matrix(t->mismatch_pear, f_mismatch_pear);
	fill_matrix(t->mismatch_rdp, f_mismatch_rdp);
	fill_matrix(t->mismatch_rdp_asm, f_mismatch_rdp_asm);
	fill_matrix(t->match_uparse, f_match_uparse);
	fill_matrix(t->mismatch_uparse, f_mismatch_uparse);
	for (int k = 0; k < PO_NQ; k++) {
		double p = phred_p(k);
		/* mktable.c:63-73: p == 1 (PHRED 0) scores -2, otherwise log(1-p); log_output=false. */
		t->score[k] = g_round((p == 1) ? -2 : log(1.0 - p));
		/* mktable.c:75-82 */
		t->score_err[k] = g_round(log(p));
	}
}

const po_tables *po_get_tables(void) {
	pthread_once(&g_tables_once, build_tables);
	return &g_tables;
}

/* ------------------------------------------------------------ small helpers */

/* prob.h:23 */
static inline int clampq(char x) {
	return x > PO_PHREDMAX ? PO_PHREDMAX : (x < 0 ? 0 : x);
}

/* pandaseq-nt.h:59 */
static inline int is_n(char nt) { return nt == (char) 0x0F; }

/* pandaseq-nt.h:55 -- popcount of the low bits != 1 */
static inline int is_degenerate(char nt) {
	return (((((unsigned int) (nt)) * 0x200040008001ULL & 0x111111111111111ULL) % 0xf) != 1);
}

void po_config_default(po_config *cfg, int algo) {
	/* assembler_support.c:36-99 defaults; algo_simple_bayes.c:110-115; algo_pear.c:103-108 */
	memset(cfg, 0, sizeof *cfg);
	cfg->algo = algo;
	cfg->minoverlap = 2;
	cfg->maxoverlap = 0;
	cfg->num_kmers = 2;
	cfg->threshold = log(0.6);
	cfg->primer_penalty = 0;
	cfg->sb_q = 0.36;
	cfg->pear_random_base = log(0.25);
}

/* ------------------------------------------------- per-algorithm callbacks */

/* algo_simple_bayes.c:126-135 */
static void sb_params(double q, double *pmatch, double *pmismatch) {
	*pmatch = log(0.25 * (1 - 2 * q + q * q));
	*pmismatch = log((3 * q - 2 * q * q) / 18.0);
}

double po_overlap_probability(const po_config *cfg, const po_qual *fwd, size_t flen,
                              const po_qual *rev, size_t rlen, size_t overlap) {
	const po_tables *t = po_get_tables();
	size_t i;
	switch (cfg->algo) {
	case PO_SIMPLE_BAYES:{
			/* algo_simple_bayes.c:33-66 */
			size_t matches = 0, mismatches = 0, unknowns = 0;
			double pmatch, pmismatch;
			sb_params(cfg->sb_q, &pmatch, &pmismatch);
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;
				char f = fwd[fi].nt, r = rev[ri].nt;
				if (is_n(f) || is_n(r))
					unknowns++;
				else if ((f & r) != 0)
					matches++;
				else
					mismatches++;
			}
			if (overlap >= flen && overlap >= rlen)
				return (t->qual_nn * unknowns + matches * pmatch + mismatches * pmismatch);
			return (t->qual_nn * (flen + rlen - 2 * overlap + unknowns) + matches * pmatch + mismatches * pmismatch);
		}
	case PO_PEAR:{
			/* algo_pear.c:32-59.  Lines 52/54 index the FORWARD read's qualities with
			 * rindex (reference quirk).  When rindex >= flen the
			 * reference reads past the array; callers of this oracle (and of the
			 * reference in our tests) pad forward[] with zeroed entries up to
			 * PO_MAX_LEN, which pins that read to quality 0. */
			double prob = 0;
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;
				char f = fwd[fi].nt, r = rev[ri].nt;
				char qb = ((size_t) ri < flen) ? fwd[ri].qual : (char) 0;
				if (is_n(f) || is_n(r))
					prob -= cfg->pear_random_base;
				else if ((f & r) != 0)
					prob += t->match_pear[clampq(fwd[fi].qual)][clampq(qb)];
				else
					prob += t->mismatch_pear[clampq(fwd[fi].qual)][clampq(qb)];
			}
			return prob;
		}
	case PO_RDP_MLE:{
			/* algo_rdp_mle.c:43-74: bounds use '>' (never true for a legal overlap). */
			double prob = 0;
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;	/* '>=': the reference's '>' would read one past the end; same guard as the others */
				int fq = clampq(fwd[fi].qual), rq = clampq(rev[ri].qual);
				if ((fwd[fi].nt & rev[ri].nt) != 0)
					prob += t->match_sb[fq][rq] - t->qual_nn;
				else
					prob += t->mismatch_rdp[fq][rq] - t->qual_nn;
			}
			return prob;
		}
	case PO_EA_UTIL:{
			/* algo_ea_util.c:29-56 */
			size_t mismatches = 0, real_overlap = 0;
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;
				char f = fwd[fi].nt, r = rev[ri].nt;
				if (is_n(f) || is_n(r) || (f & r) == 0)
					mismatches++;
				real_overlap++;
			}
			return log((((double) mismatches) * mismatches + 1) / real_overlap);
		}
	case PO_STITCH:{
			/* algo_stitch.c:28-56: the score is a size_t, so a net-negative score wraps around */
			size_t score = 0;
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;
				char f = fwd[fi].nt, r = rev[ri].nt;
				if (is_n(f) || is_n(r))
					score += 0;
				else if ((f & r) != 0)
					score += 1;
				else
					score -= 1;
			}
			return log(score / (double) (flen + rlen));
		}
	case PO_UPARSE:{
			/* algo_uparse.c:33-66 with the parameters of algo_uparse.c:126-135 */
			size_t matches = 0, mismatches = 0, unknowns = 0;
			double q = cfg->sb_q;
			double pmatch = log(1 - q * q * (1 - 2 * q + 4 * q * q / 3));
			double pmismatch = log(1 - 4 * q / 3 / (2 * q - 4 * q * q / 3));
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;
				char f = fwd[fi].nt, r = rev[ri].nt;
				if (is_n(f) || is_n(r))
					unknowns++;
				else if ((f & r) != 0)
					matches++;
				else
					mismatches++;
			}
			if (overlap >= flen && overlap >= rlen)
				return (t->qual_nn * unknowns + matches * pmatch + mismatches * pmismatch);
			return (t->qual_nn * (flen + rlen - 2 * overlap + unknowns) + matches * pmatch + mismatches * pmismatch);
		}
	case PO_FLASH:{
			/* algo_flash.c:30-60: size_t division, so log(0) unless every base mismatches. */
			size_t mismatches = 0, real_overlap = 0;
			for (i = 0; i < overlap; i++) {
				int fi = (int) (flen + i - overlap);
				int ri = (int) (rlen - i - 1);
				if (fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
					continue;
				char f = fwd[fi].nt, r = rev[ri].nt;
				if (is_n(f) || is_n(r) || (f & r) == 0)
					mismatches++;
				real_overlap++;
			}
			return real_overlap == 0 ? -2 : log((double) (mismatches / real_overlap));
		}
	}
	return -INFINITY;
}

double po_match_probability(const po_config *cfg, int match, char a, char b) {
	const po_tables *t = po_get_tables();
	switch (cfg->algo) {
	case PO_SIMPLE_BAYES:	/* algo_simple_bayes.c:68-75 */
		return (match ? t->match_sb : t->mismatch_sb)[clampq(a)][clampq(b)];
	case PO_PEAR:		/* algo_pear.c:61-68 */
		return (match ? t->match_pear : t->mismatch_pear)[clampq(a)][clampq(b)];
	case PO_RDP_MLE:	/* algo_rdp_mle.c:29-41 */
		if (match) {
			char hi = (a >= b) ? a : b;
			return t->score[clampq(hi)];
		}
		return t->mismatch_rdp_asm[clampq(a)][clampq(b)];
	case PO_EA_UTIL:	/* algo_ea_util.c:58-67: the higher quality, match or not */
		return t->score[(a > b) ? clampq(a) : clampq(b)];
	case PO_STITCH:		/* algo_stitch.c:58-66 */
		return (match ? t->match_sb : t->mismatch_sb)[clampq(a)][clampq(b)];
	case PO_UPARSE:		/* algo_uparse.c:68-75 */
		return (match ? t->match_uparse : t->mismatch_uparse)[clampq(a)][clampq(b)];
	case PO_FLASH:{	/* algo_flash.c:62-80 */
			int s;
			if (match) {
				s = (a > b) ? clampq(a) : clampq(b);
			} else {
				s = clampq(a) - clampq(b);
				if (s < 0)
					s = -s;
				if (s < 2)
					s = 2;
			}
			return t->score[s];
		}
	}
	return 0;
}

/* ----------------------------------------------------------- primer locate */

/* offset.c:47-101 with the qual_base_score scorer (offset.c:92-101). */
size_t po_compute_offset_qual(double threshold, double penalty, int reverse,
                              const po_qual *hay, size_t hay_len, const char *needle, size_t needle_len) {
	const po_tables *t = po_get_tables();
	double ring[PO_MAX_LEN];
	double best = exp(needle_len * threshold);
	size_t best_index = 0;
	if (needle_len > hay_len || needle_len == 0 || needle_len > PO_MAX_LEN)
		return 0;
	for (size_t k = 0; k < needle_len; k++)
		ring[k] = -INFINITY;
	for (size_t index = 0; index < hay_len; index++) {
		size_t slot = index % needle_len;
		double last = exp(ring[slot] / (index + 1)) - index * penalty;
		if (last > best) {
			best = last;
			best_index = index + 1;
		}
		ring[slot] = 0;
		const po_qual *base = &hay[reverse ? (hay_len - index - 1) : index];
		int phred = clampq(base->qual);
		for (ptrdiff_t x = (ptrdiff_t) (needle_len > index ? index : needle_len - 1); x >= 0; x--) {
			if (!is_n(needle[x])) {
				size_t dst = (index - (size_t) x) % needle_len;
				ring[dst] += ((base->nt & needle[x]) != 0) ? t->score[phred] : t->score_err[phred];
			}
		}
	}
	return best_index;
}

/* offset.c:35-38 */
static double log1mexp(double p) {
	return (p > 0.69314718055994530942) ? log1p(-exp(-p)) : log(-expm1(-p));
}

/* offset.c:47-90 with the result_base_score scorer (offset.c:114-133): the haystack is the assembled sequence.
 * For a log-probability p < 0 the "not p" branch evaluates log(-expm1(-p)) = log(negative) = NaN, so any
 * mismatching base poisons that start offset; reproduced literally through libm. */
static size_t compute_offset_result(double threshold, double penalty, int reverse, const char *nt, const double *p,
                                    size_t hay_len, const char *needle, size_t needle_len) {
	double ring[PO_MAX_LEN];
	double best = exp(needle_len * threshold);
	size_t best_index = 0;
	if (needle_len > hay_len || needle_len == 0 || needle_len > PO_MAX_LEN)
		return 0;
	for (size_t k = 0; k < needle_len; k++)
		ring[k] = -INFINITY;
	for (size_t index = 0; index < hay_len; index++) {
		size_t slot = index % needle_len;
		double last = exp(ring[slot] / (index + 1)) - index * penalty;
		if (last > best) {
			best = last;
			best_index = index + 1;
		}
		ring[slot] = 0;
		size_t el = reverse ? (hay_len - index - 1) : index;
		double pr = p[el], notp = log1mexp(pr);
		for (ptrdiff_t x = (ptrdiff_t) (needle_len > index ? index : needle_len - 1); x >= 0; x--) {
			if (!is_n(needle[x])) {
				size_t dst = (index - (size_t) x) % needle_len;
				ring[dst] += ((nt[el] & needle[x]) != 0) ? pr : notp;
			}
		}
	}
	return best_index;
}

/* ------------------------------------------------------------------- align */

typedef struct {
	int status;
	int slow;
	int overlap;
	int seq_len;
	int mismatches;
	int degenerates;
	int examined;
	int fwd_offset;
	int rev_offset;
	double quality;
	double est_prob;
	char nt[2 * PO_MAX_LEN];
	double p[2 * PO_MAX_LEN];
} po_one;

/* misc.h:41: 2-bit code of one base for the rolling 8-mer. */
static inline unsigned kcode(char nt) {
	return nt == 8 ? 3u : nt == 4 ? 2u : nt == 2 ? 1u : 0u;
}

/* K1-K3 of align() (assembler.c:92-116): the candidate-overlap bit list before ALL_BITS_IF_NONE.  `table` is the
 * 65536 x 2 position table, all zero on entry and on exit; `bits` holds nbits / 32 + 1 zeroed words. */
static void seed_bits(uint16_t *table, const po_qual *F, size_t flen, const po_qual *R, size_t rlen, size_t mo,
                      size_t nbits, uint32_t *bits) {
	/* K1: forward k-mers, assembler.c:92-101 + misc.h:41-42 */
	{
		unsigned code = 0;
		int bad = 8;
		for (size_t p = 0; p < flen; p++) {
			code = ((code << 2) | kcode(F[p].nt)) & 0xFFFFu;
			if (is_n(F[p].nt)) {
				bad = 8;
			} else if (bad > 0) {
				bad--;
			} else {
				uint16_t *slot = &table[code * 2];
				if (slot[0] == 0)
					slot[0] = (uint16_t) p;
				else if (slot[1] == 0)
					slot[1] = (uint16_t) p;
				/* else: lost k-mer (assembler.c:95-97) */
			}
		}
	}
	/* K2: reverse k-mers walked from the end, assembler.c:104-110 + misc.h:43 */
	{
		unsigned code = 0;
		int bad = 8;
		for (ptrdiff_t pr = (ptrdiff_t) rlen - 1; pr >= 0; pr--) {
			code = ((code << 2) | kcode(R[pr].nt)) & 0xFFFFu;
			if (is_n(R[pr].nt)) {
				bad = 8;
			} else if (bad > 0) {
				bad--;
			} else {
				const uint16_t *slot = &table[code * 2];
				for (int j = 0; j < 2 && slot[j] != 0; j++) {
					int index = (int) (flen + rlen - (size_t) pr - slot[j] - mo - 1);
					if (index >= 0 && (size_t) index < nbits)
						bits[index / 32] |= (1u << (index % 32));
				}
			}
		}
	}
	/* K3: clear, assembler.c:113-116 */
	{
		unsigned code = 0;
		int bad = 8;
		for (size_t p = 0; p < flen; p++) {
			code = ((code << 2) | kcode(F[p].nt)) & 0xFFFFu;
			if (is_n(F[p].nt)) {
				bad = 8;
			} else if (bad > 0) {
				bad--;
			} else {
				table[code * 2] = 0;
				table[code * 2 + 1] = 0;
			}
		}
	}
}

/* assembler.c:48-250.  `table` is the 65536 x 2 position table, all zero on entry and on exit. */
static int align_pair(const po_config *cfg, uint16_t *table, const po_qual *F, size_t flen,
                      const po_qual *R, size_t rlen, size_t fo, size_t ro, po_one *out) {
	const po_tables *t = po_get_tables();
	const double qual_nn = t->qual_nn;
	const size_t mo = (size_t) cfg->minoverlap;
	size_t maxov = flen + rlen - mo - fo - ro - 1;	/* assembler.c:59 */
	double best = qual_nn * (flen + rlen);	/* assembler.c:60 */
	ptrdiff_t bestov = -1;

	if (mo + fo >= flen || mo + ro >= rlen)	/* assembler.c:73-76 */
		return 0;
	if (cfg->maxoverlap == 0)	/* assembler.c:78-82 */
		maxov = flen < rlen ? flen : rlen;
	else if (maxov > (size_t) cfg->maxoverlap)
		maxov = (size_t) cfg->maxoverlap;

	size_t nbits = mo <= maxov ? (maxov - mo + 1) : 1;	/* assembler.c:84 */
	uint32_t bits[(2 * PO_MAX_LEN) / 32 + 2];
	size_t nwords = nbits / 32 + 1;
	memset(bits, 0, nwords * sizeof(uint32_t));

	seed_bits(table, F, flen, R, rlen, mo, nbits, bits);
	/* assembler.c:118 */
	{
		uint32_t any = 0;
		for (size_t w = 0; w < nwords; w++)
			any |= bits[w];
		if (any == 0)
			memset(bits, 0xFF, nwords * sizeof(uint32_t));
	}
	/* K4: assembler.c:120-133 */
	out->examined = 0;
	for (size_t c = 0; c < nbits; c++) {
		if (!(bits[c / 32] & (1u << (c % 32))))
			continue;
		size_t ov = c + mo;
		double pr = po_overlap_probability(cfg, F, flen, R, rlen, ov);
		if (pr > best && ov >= mo) {
			best = pr;
			bestov = (ptrdiff_t) ov;
		}
		out->examined++;
	}
	if ((size_t) out->examined == maxov - mo + 1)	/* assembler.c:135-137 */
		out->slow = 1;
	if (bestov == -1)
		return 0;

	/* K6: assembler.c:145-250 */
	ptrdiff_t len = (ptrdiff_t) flen - (ptrdiff_t) fo - bestov + (ptrdiff_t) rlen - (ptrdiff_t) ro + 1;
	if (len <= 0)
		return 0;
	if ((size_t) len > 2 * PO_MAX_LEN)
		return 0;
	out->seq_len = (int) (len - 1);
	out->degenerates = 0;
	ptrdiff_t df = (ptrdiff_t) flen - (ptrdiff_t) fo - bestov;
	ptrdiff_t dr = (ptrdiff_t) rlen - (ptrdiff_t) ro - bestov;
	ptrdiff_t dfp = df < 0 ? 0 : df, dfn = df > 0 ? 0 : df;
	ptrdiff_t drp = dr < 0 ? 0 : dr, drn = dr > 0 ? 0 : dr;
	double fquality = 0, oquality = 0, rquality = 0;

	for (ptrdiff_t i = 0; i < dfp; i++) {	/* assembler.c:162-173 */
		size_t fi = (size_t) i + fo;
		double q = t->score[clampq(F[fi].qual)];
		out->nt[i] = F[fi].nt;
		out->p[i] = q;
		if (is_degenerate(F[fi].nt))
			out->degenerates++;
		fquality += q;
	}
	size_t unmasked_f = flen, unmasked_r = rlen;	/* assembler.c:176-177 */
	while (unmasked_f > 0 && F[unmasked_f - 1].qual == (char) 2)
		unmasked_f--;
	while (unmasked_r > 0 && R[unmasked_r - 1].qual == (char) 2)
		unmasked_r--;

	out->mismatches = 0;
	ptrdiff_t nover = bestov + dfn + drn;	/* assembler.c:181 */
	for (ptrdiff_t i = 0; i < nover; i++) {
		ptrdiff_t index = dfp + i;
		ptrdiff_t fi = (ptrdiff_t) fo + dfp + i;
		ptrdiff_t ri = (ptrdiff_t) rlen - i - 1 + dfn;
		if (index < 0 || fi < 0 || ri < 0 || (size_t) fi >= flen || (size_t) ri >= rlen)
			continue;	/* assembler.c:191 (checked before the read here) */
		char fn = F[fi].nt, rn = R[ri].nt;
		char fq = F[fi].qual, rq = R[ri].qual;
		int ismatch = (rn & fn) != 0;
		int fmasked = (size_t) fi >= unmasked_f, rmasked = (size_t) ri >= unmasked_r;
		double q;
		char nt;
		if (!ismatch)
			out->mismatches++;
		if (fmasked && rmasked)
			q = qual_nn;
		else if (fmasked)
			q = t->score[clampq(rq)];
		else if (rmasked)
			q = t->score[clampq(fq)];
		else
			q = po_match_probability(cfg, ismatch, fq, rq);
		if (ismatch)
			nt = (char) (rn & fn);
		else
			nt = (fq < rq) ? rn : fn;	/* assembler.c:215-219 */
		out->nt[index] = nt;
		out->p[index] = q;
		if (is_degenerate(nt))
			out->degenerates++;
		oquality += q;
	}
	for (ptrdiff_t i = 0; i < drp; i++) {	/* assembler.c:231-243 */
		ptrdiff_t index = df + bestov + i;
		ptrdiff_t ri = (ptrdiff_t) rlen - bestov - i - 1;
		double q = t->score[clampq(R[ri].qual)];
		rquality += q;
		out->nt[index] = R[ri].nt;
		out->p[index] = q;
		if (is_degenerate(R[ri].nt))
			out->degenerates++;
	}
	out->quality = (fquality + rquality + oquality) / len;	/* assembler.c:244 */
	out->overlap = (int) bestov;
	out->est_prob = best;
	return 1;
}

/* assembler.c:252-348 (module hooks are host callbacks outside this path). */
static void assemble_pair(const po_config *cfg, uint16_t *table, const po_qual *F, size_t flen,
                          const po_qual *R, size_t rlen, po_one *out) {
	size_t fo, ro;
	memset(out, 0, offsetof(po_one, nt));
	if (flen < 2 || rlen < 2) {
		out->status = PO_BADR;
		return;
	}
	if (!cfg->post_primers) {
		if (cfg->forward_primer_length > 0) {
			fo = po_compute_offset_qual(cfg->threshold, cfg->primer_penalty, 0, F, flen, cfg->forward_primer, (size_t) cfg->forward_primer_length);
			if (fo == 0) {
				out->status = PO_NOFP;
				return;
			}
			fo--;
		} else {
			fo = (size_t) cfg->forward_trim;
		}
		out->fwd_offset = (int) fo;
		if (cfg->reverse_primer_length > 0) {
			ro = po_compute_offset_qual(cfg->threshold, cfg->primer_penalty, 0, R, rlen, cfg->reverse_primer, (size_t) cfg->reverse_primer_length);
			if (ro == 0) {
				out->status = PO_NORP;
				return;
			}
			ro--;
		} else {
			ro = (size_t) cfg->reverse_trim;
		}
		out->rev_offset = (int) ro;
	} else {
		fo = ro = 0;		/* assembler.c:285-288 */
	}
	if ((flen < rlen ? flen : rlen) < (size_t) cfg->minoverlap) {
		out->status = PO_BADR;
		return;
	}
	if (!align_pair(cfg, table, F, flen, R, rlen, fo, ro, out)) {
		out->status = PO_NOALGN;
		return;
	}
	if (cfg->post_primers) {	/* assembler.c:300-333 */
		if (cfg->forward_primer_length > 0) {
			fo = compute_offset_result(cfg->threshold, cfg->primer_penalty, 0, out->nt, out->p, (size_t) out->seq_len, cfg->forward_primer, (size_t) cfg->forward_primer_length);
			if (fo == 0) {
				out->status = PO_NOFP;
				return;
			}
			fo--;
		} else {
			fo = (size_t) cfg->forward_trim;
		}
		out->fwd_offset = (int) fo;
		if (cfg->reverse_primer_length > 0) {
			ro = compute_offset_result(cfg->threshold, cfg->primer_penalty, 1, out->nt, out->p, (size_t) out->seq_len, cfg->reverse_primer, (size_t) cfg->reverse_primer_length);
			if (ro == 0) {
				out->status = PO_NORP;
				return;
			}
			ro--;
		} else {
			ro = (size_t) cfg->reverse_trim;
		}
		out->rev_offset = (int) ro;
		if ((size_t) out->seq_len <= fo + ro) {
			out->status = PO_NOFP;	/* sic: assembler.c:324-328 counts this as a missing forward primer */
			return;
		}
		out->seq_len -= (int) (fo + ro);
		memmove(out->nt, out->nt + fo, (size_t) out->seq_len);
		memmove(out->p, out->p + fo, (size_t) out->seq_len * sizeof(double));
	}
	if (out->quality < cfg->threshold) {
		out->status = PO_LOWQ;
		return;
	}
	out->status = PO_OK;
}

/* ------------------------------------------------------------ batch driver */

/* hang.c:39-72 with the reversed copies of hang.c:103-106.  Returns 0 when the pair is dropped. */
static int trim_overhangs(const po_config *cfg, const po_qual *F, size_t *flen, const po_qual *R, size_t *rlen) {
	char rev[PO_MAX_LEN];
	if (cfg->hang_forward_length > 0) {
		size_t n = (size_t) cfg->hang_forward_length, off;
		for (size_t k = 0; k < n; k++)
			rev[n - k - 1] = cfg->hang_forward[k];
		off = po_compute_offset_qual(cfg->hang_threshold, 0, 1, F, *flen, rev, n);
		if (off == 0) {
			if (!cfg->hang_skip)
				return 0;
		} else {
			*flen -= off - 1;
		}
	}
	if (cfg->hang_reverse_length > 0) {
		size_t n = (size_t) cfg->hang_reverse_length, off;
		for (size_t k = 0; k < n; k++)
			rev[n - k - 1] = cfg->hang_reverse[k];
		off = po_compute_offset_qual(cfg->hang_threshold, 0, 1, R, *rlen, rev, n);
		if (off == 0) {
			if (!cfg->hang_skip)
				return 0;
		} else {
			*rlen -= off - 1;
		}
	}
	return 1;
}

/* plugin_pear_test.c:18-39.  The reference converts ceil(..) - 1 to size_t; a negative (or NaN) value makes its inner loop
 * run for 2^64 iterations -- it never comes back -- while adding nothing once k exceeds i (lgamma of a non-positive integer
 * is +inf, the term exp(-inf) is 0).  Defined here as what that loop has accumulated by then: all i + 1 terms. */
static int pear_test_check(double alpha, double beta, double cutoff, size_t overlap, size_t mismatches, size_t flen, size_t rlen) {
	double product = 1;
	double oes = alpha * (overlap - mismatches) + beta * mismatches;
	for (size_t i = overlap; i < flen && i < rlen; i++) {
		double sum = 0;
		double lraw = ceil((oes - beta * i) / (alpha - beta)) - 1;
		size_t l_i = (lraw >= 0 && lraw < (double) (i + 1)) ? (size_t) lraw : i + 1;
		for (size_t k = 0; k < l_i; k++) {
			double i_choose_k = lgamma(i + 1) - lgamma(k + 1) - lgamma(i - k + 1);
			sum += exp(i_choose_k + k * log(0.25) + (i - k) * log(0.75));
		}
		product *= sum;
	}
	return cutoff > 1 - product * product;
}

/* args_assembler.c:106-115,233-239,268-275; plugin_min_overlapbits.c:17-23; plugin_completely_miss_the_point.c:9-16;
 * plugin_min_phred.c:8-22; plugin_pear_test.c.  Returns the index of the first check that fails, -1 if all pass. */
static int first_failing_filter(const po_config *cfg, const po_one *one, size_t flen, size_t rlen) {
	for (int k = 0; k < cfg->nfilters && k < 7; k++) {
		const struct po_filter *f = &cfg->filters[k];
		int pass = 1;
		switch (f->kind) {
		case PO_FILTER_NO_N: pass = one->degenerates == 0; break;
		case PO_FILTER_SHORT: pass = (size_t) one->seq_len >= (size_t) f->ivalue; break;
		case PO_FILTER_LONG: pass = (size_t) one->seq_len <= (size_t) f->ivalue; break;
		case PO_FILTER_MIN_OVERLAPBITS: pass = f->dvalue * M_LN2 <= one->est_prob; break;	/* the plugin takes bits and compares nats */
		case PO_FILTER_MISS_THE_POINT: pass = (size_t) one->mismatches <= (size_t) f->ivalue; break;
		case PO_FILTER_MIN_PHRED:
			for (int it = 0; it < one->seq_len && pass; it++)
				if (po_result_phred(one->p[it]) < f->ivalue)
					pass = 0;
			break;
		case PO_FILTER_PEAR_TEST:
			pass = pear_test_check(f->dvalue, f->dvalue2, f->dvalue3, (size_t) one->overlap, (size_t) one->mismatches, flen, rlen);
			break;
		}
		if (!pass)
			return k;
	}
	return -1;
}

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
} po_job;

static void *run_job(void *arg) {
	po_job *job = arg;
	po_flat_out *o = job->out;
	uint16_t *table = calloc(65536 * 2, sizeof(uint16_t));
	po_one *one = malloc(sizeof(po_one));
	/* pear's forward[rindex] quirk needs a padded forward buffer (see po_overlap_probability). */
	po_qual fpad[PO_MAX_LEN];
	if (table == NULL || one == NULL) {
		job->failed = 1;
		free(table);
		free(one);
		return NULL;
	}
	memset(job->counters, 0, sizeof job->counters);
	for (size_t i = job->begin; i < job->end; i++) {
		size_t flen = job->f_off[i + 1] - job->f_off[i];
		size_t rlen = job->r_off[i + 1] - job->r_off[i];
		const po_qual *F = job->f_data + job->f_off[i];
		const po_qual *R = job->r_data + job->r_off[i];
		if (flen > PO_MAX_LEN || rlen > PO_MAX_LEN) {
			job->failed = 1;
			break;
		}
		/* hang.c:39-72: the overhang trimmer sits between the reader and the assembler */
		if (!trim_overhangs(job->cfg, F, &flen, R, &rlen)) {
			if (o->status) o->status[i] = PO_SKIP;
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
			continue;
		}
		if (job->cfg->algo == PO_PEAR) {
			memset(fpad, 0, sizeof fpad);
			memcpy(fpad, F, flen * sizeof(po_qual));
			F = fpad;
		}
		assemble_pair(job->cfg, table, F, flen, R, rlen, one);
		if (one->status == PO_OK) {	/* module_checkseq, module.c:124-137: the first failing check rejects */
			int k = first_failing_filter(job->cfg, one, flen, rlen);      /* the lengths the assembler saw (after the overhang trimmer) */
			if (k >= 0) {
				one->status = PO_FILTERED + k;
				job->counters[PO_C_REJECTED + k]++;
			}
		}
		job->counters[PO_C_COUNT]++;
		if (one->slow)
			job->counters[PO_C_SLOW]++;
		switch (one->status) {
		case PO_OK:
			job->counters[PO_C_OK]++;
			job->counters[PO_C_OVERLAPS + one->overlap]++;
			if (job->counters[PO_C_LONGEST] < one->overlap)
				job->counters[PO_C_LONGEST] = one->overlap;
			break;
		case PO_BADR: job->counters[PO_C_BADR]++; break;
		case PO_NOFP: job->counters[PO_C_NOFP]++; break;
		case PO_NORP: job->counters[PO_C_NORP]++; break;
		case PO_NOALGN: job->counters[PO_C_NOALGN]++; break;
		case PO_LOWQ: job->counters[PO_C_LOWQ]++; break;
		}
		if (o->status) o->status[i] = (uint8_t) one->status;
		if (o->slow) o->slow[i] = (uint8_t) one->slow;
		int emitted = (one->status == PO_OK || one->status == PO_LOWQ || one->status >= PO_FILTERED);
		if (o->overlap) o->overlap[i] = emitted ? one->overlap : 0;
		if (o->seq_len) o->seq_len[i] = emitted ? one->seq_len : 0;
		if (o->mismatches) o->mismatches[i] = emitted ? one->mismatches : 0;
		if (o->degenerates) o->degenerates[i] = emitted ? one->degenerates : 0;
		if (o->examined) o->examined[i] = one->examined;
		if (o->fwd_offset) o->fwd_offset[i] = one->fwd_offset;
		if (o->rev_offset) o->rev_offset[i] = one->rev_offset;
		if (o->quality) o->quality[i] = emitted ? one->quality : 0;
		if (o->est_prob) o->est_prob[i] = emitted ? one->est_prob : 0;
		if (emitted && o->seq_stride >= one->seq_len) {
			if (o->seq_nt) {
				uint8_t *dst = o->seq_nt + i * (size_t) o->seq_stride;
				memcpy(dst, one->nt, (size_t) one->seq_len);
				memset(dst + one->seq_len, 0, (size_t) (o->seq_stride - one->seq_len));
			}
			if (o->seq_p) {
				double *dst = o->seq_p + i * (size_t) o->seq_stride;
				memcpy(dst, one->p, (size_t) one->seq_len * sizeof(double));
				memset(dst + one->seq_len, 0, (size_t) (o->seq_stride - one->seq_len) * sizeof(double));
			}
		} else {
			if (o->seq_nt) memset(o->seq_nt + i * (size_t) o->seq_stride, 0, (size_t) o->seq_stride);
			if (o->seq_p) memset(o->seq_p + i * (size_t) o->seq_stride, 0, (size_t) o->seq_stride * sizeof(double));
		}
	}
	free(table);
	free(one);
	return NULL;
}

int po_assemble_flat(const po_config *cfg, size_t n,
                     const po_qual *f_data, const uint64_t *f_off,
                     const po_qual *r_data, const uint64_t *r_off,
                     po_flat_out *out, int threads) {
	if (cfg->num_kmers != 2 || cfg->minoverlap < 2
	    || cfg->algo < PO_SIMPLE_BAYES || cfg->algo > PO_UPARSE)
		return -1;
	if (threads < 1)
		threads = 1;
	if ((size_t) threads > n)
		threads = n ? (int) n : 1;
	po_job *jobs = calloc((size_t) threads, sizeof(po_job));
	pthread_t *tids = calloc((size_t) threads, sizeof(pthread_t));
	(void) po_get_tables();
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
		for (int k = 0; k < threads; k++) {
			for (int c = 0; c < PO_NCOUNTERS; c++) {
				if (c == PO_C_LONGEST) {
					if (out->counters[c] < jobs[k].counters[c])
						out->counters[c] = jobs[k].counters[c];
				} else {
					out->counters[c] += jobs[k].counters[c];
				}
			}
		}
	}
	free(jobs);
	free(tids);
	return failed ? -1 : 0;
}

/* Test hook: the candidate-overlap bit list of one pair exactly as K1-K3 leave it (assembler.c:84-116, before
 * ALL_BITS_IF_NONE), with forward_offset = reverse_offset = 0.  Returns the number of bits (0: align() fails before
 * seeding, assembler.c:73-76); `bits` must hold (2 * PO_MAX_LEN) / 32 + 2 words. */
size_t po_seed_bits(const po_config *cfg, const po_qual *F, size_t flen, const po_qual *R, size_t rlen, uint32_t *bits) {
	static __thread uint16_t *table;
	const size_t mo = (size_t) cfg->minoverlap;
	size_t maxov = flen + rlen - mo - 1;
	if (table == NULL)
		table = calloc(65536 * 2, sizeof(uint16_t));
	if (mo >= flen || mo >= rlen)
		return 0;
	if (cfg->maxoverlap == 0)
		maxov = flen < rlen ? flen : rlen;
	else if (maxov > (size_t) cfg->maxoverlap)
		maxov = (size_t) cfg->maxoverlap;
	const size_t nbits = mo <= maxov ? (maxov - mo + 1) : 1;
	memset(bits, 0, (nbits / 32 + 1) * sizeof(uint32_t));
	seed_bits(table, F, flen, R, rlen, mo, nbits, bits);
	return nbits;
}
