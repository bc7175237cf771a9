/* pb_internal.h -- declarations shared by the host C files and the CUDA file.
 * Not installed; the public ABI is include/pandaseq_b200.h. */
#ifndef PB_INTERNAL_H
#define PB_INTERNAL_H
#include "pandaseq_b200.h"
#include <pthread.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_NQ (PB_PHREDMAX + 1)   /* 47 PHRED values */
#define PB_NQM (PB_NQ + 1)        /* + index 47 = "masked / absent read" */

/* What the kernels read, built on the host from a pb_config (pb_luts.c) and kept in HBM.
 *  recon[k][a][b]  : per-base log p during reconstruction (assembler.c:162-243), k = bases match;
 *                    a/b = clamped PHRED of the forward/reverse base, or 47 when that read is absent
 *                    (forward-only / reverse-only stretch) or B-cliff masked (assembler.c:176-210).
 *  over[k][a][b]   : per-base term of the overlap score for rdp_mle (algo_rdp_mle.c:68-70); pear's terms
 *                    (algo_pear.c:52-54) equal recon[k][a][b] and are read from there.
 */
typedef struct {
	double recon[2][PB_NQM][PB_NQM];
	double over[2][PB_NQ][PB_NQ];
	double score[PB_NQM];       /* qual_score, [47] unused (0) */
	double score_err[PB_NQM];   /* qual_score_err */
	double qual_nn;
	double sb_pmatch;           /* algo_simple_bayes.c:132-133 */
	double sb_pmismatch;
	double pear_random_base;
	double threshold;
	double primer_penalty;
	int32_t algo;
	int32_t minoverlap;
	int32_t maxoverlap;
	int32_t forward_trim;
	int32_t reverse_trim;
	int32_t forward_primer_length;
	int32_t reverse_primer_length;
	int32_t post_primers;       /* assembler.c:300-333: locate the primers on the assembled sequence instead of the reads */
	uint8_t forward_primer[PB_MAX_LEN + 2];
	uint8_t reverse_primer[PB_MAX_LEN + 2];
	/* overhang trimmer: the two sequences REVERSED, as hang.c:103-106 stores them for its end-first scans */
	int32_t hang_forward_length;
	int32_t hang_reverse_length;
	int32_t hang_skip;
	int32_t nfilters;
	int32_t need_stage;         /* a filter reads the per-base log p (min_phred): the sequence is staged as for primers-after */
	int32_t pad0;
	double hang_threshold;
	uint8_t hang_forward[PB_MAX_LEN + 2];
	uint8_t hang_reverse[PB_MAX_LEN + 2];
	struct pb_filter filters[PB_MAX_FILTERS];
	/* pear_test: cdf[i * PB_PEAR_COLS + l] = sum over k < l of C(i,k) 0.25^k 0.75^(i-k), added in k order as
	 * plugin_pear_test.c:31-35 adds them (device pointer, set when a pear_test filter is configured) */
	const double *pear_cdf;
} pb_device_params;
#define PB_PEAR_ROWS PB_MAX_LEN
#define PB_PEAR_COLS (PB_MAX_LEN + 2)
void pb_build_pear_cdf(double *out);   /* PB_PEAR_ROWS x PB_PEAR_COLS doubles */

/* pb_luts.c */
pb_status pb_build_device_params(const pb_config *cfg, pb_device_params *out);
double pb_host_match_probability(int algo, bool match, char a, char b);
void pb_set_error(const char *fmt, ...);

/* pb_algorithm.c */
struct panda_algorithm {
	PandaAlgorithmClass clazz;
	volatile size_t refcnt;
	pthread_mutex_t mutex;
	/* private data of clazz->data_size bytes follows; aligned like a pointer, as in algo.h:27-34 */
	void *end;
};
int pb_algorithm_fill_config(PandaAlgorithm algo, pb_config *cfg);

/* per-warp global scratch for the primers-after path: 912 f64 log-probabilities + 512 B of 4-bit bases */
#define PB_SCRATCH_STRIDE (912 * 8 + 512)

/* pb_device.cu */
pb_status pb_shared_context(pb_context **out);
pb_status pb_device_context(int device, pb_context **out);
void pb_bind_thread_near_device(int device);   /* the calling thread onto the host cores next to a GPU (sysfs local_cpulist); best effort */
   /* the process-wide context of one GPU (panda_run_pool's workers) */

#ifdef __cplusplus
}
#endif
#endif
