/* pandaseq_b200.h -- C ABI of libpandaseq_b200.so
 *
 * A Blackwell (sm_100a) implementation of ONE path of PANDAseq: what happens
 * behind panda_assembler_assemble() -- primer location, k-mer seeded overlap
 * selection, and reconstruction of the merged read with posterior qualities,
 * for the simple_bayesian / pear / rdp_mle / flash scoring algorithms.
 *
 * Two layers, both plain C (no C++/torch types cross this boundary):
 *
 *   1. panda_*  -- the reference's own object API for this path, same names,
 *      argument meaning and error behaviour, so a program written against
 *      <pandaseq.h> links against this library for assembling pairs.  Each
 *      declaration cites the reference declaration/definition it replaces.
 *      The structs below are ABI-compatible with the reference's (they have to
 *      be: callers allocate panda_qual arrays and read panda_result_seq).
 *
 *   2. pb_*  -- the batch layer the panda_* calls sit on: a device context, a
 *      flat batch format, and launch entry points that take either host
 *      buffers (copies inside) or device pointers (for callers that already
 *      keep reads in HBM, e.g. bench.py via torch allocations).
 *
 * There is no CPU fallback: every entry point that computes needs a CUDA device
 * and returns an error (NULL / negative pb_status) without one.
 */
#ifndef PANDASEQ_B200_H
#define PANDASEQ_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================
 * Shared data types (ABI-compatible with the reference)
 * ====================================================================== */

/* pandaseq-common.h:176 -- 4-bit one-hot IUPAC code in a char: A=1 C=2 G=4 T=8,
 * degenerate = OR, N = 15, 0 = invalid (pandaseq-nt.h:35-51). */
typedef char panda_nt;
#define PANDA_NT_Z ((panda_nt) 0)
#define PANDA_NT_A ((panda_nt) 1)
#define PANDA_NT_C ((panda_nt) 2)
#define PANDA_NT_G ((panda_nt) 4)
#define PANDA_NT_T ((panda_nt) 8)
#define PANDA_NT_N ((panda_nt) 15)

/* pandaseq-common.h:210-219 */
typedef struct {
	panda_nt nt;
	char qual;
} panda_qual;

/* pandaseq-common.h:224-233 */
typedef struct {
	panda_nt nt;
	double p; /* log probability that the base is right */
} panda_result;

#define PANDA_TAG_LEN 50
/* pandaseq-common.h:238-247 (368 bytes; carried through, never interpreted here) */
typedef struct {
	char instrument[100];
	char run[100];
	char flowcell[100];
	int lane;
	int tile;
	int x;
	int y;
	char tag[PANDA_TAG_LEN];
} panda_seq_identifier;

/* pandaseq-common.h:277-330 */
typedef struct {
	double quality;
	size_t degenerates;
	panda_seq_identifier name;
	panda_result *sequence;
	size_t sequence_length;
	panda_qual const *forward;
	size_t forward_length;
	panda_qual const *reverse;
	size_t reverse_length;
	size_t forward_offset;
	size_t reverse_offset;
	size_t overlap_mismatches;
	size_t overlaps_examined;
	size_t overlap;
	double estimated_overlap_probability;
} panda_result_seq;

typedef struct panda_algorithm *PandaAlgorithm;
typedef const struct panda_algorithm_class *PandaAlgorithmClass;
typedef struct panda_assembler *PandaAssembler;
typedef struct panda_log_proxy *PandaLogProxy; /* accepted and ignored: logging is out of scope */
typedef struct panda_mux *PandaMux;            /* accepted and ignored by panda_run_pool: the mux is out of scope */

/* pandaseq-common.h:335-340, 366-397, 402-403 */
typedef PandaAlgorithm (*PandaAlgorithmCreate) (const char *args);
typedef double (*PandaComputeMatch) (void *private_data, bool match, char a, char b);
typedef double (*PandaComputeOverlap) (void *private_data, const panda_qual *forward, size_t forward_length,
                                       const panda_qual *reverse, size_t reverse_length, size_t overlap);
typedef void (*PandaDestroy) (void *user_data);
/* pandaseq-common.h:481-497 -- pull source; arrays valid until the next call */
typedef bool (*PandaNextSeq) (panda_seq_identifier *id, const panda_qual **forward, size_t *forward_length,
                              const panda_qual **reverse, size_t *reverse_length, void *user_data);
/* pandaseq-common.h:499-507 */
typedef bool (*PandaOutputSeq) (const panda_result_seq *sequence, void *user_data);
/* pandaseq-common.h:408-426 */
typedef void (*PandaFailAlign) (PandaAssembler assembler, const panda_seq_identifier *id,
                                const panda_qual *forward, size_t forward_length,
                                const panda_qual *reverse, size_t reverse_length, void *user_data);
/* pandaseq-common.h:356-363, 509-522 -- the checks of a module: HOST functions.  The batch driver runs the pre-checks while it
 * stages a pair (a rejected pair never reaches the device) and the checks while it hands the assembled pairs out. */
typedef struct panda_module *PandaModule;
typedef bool (*PandaCheck) (PandaLogProxy logger, const panda_result_seq *sequence, void *user_data);
typedef bool (*PandaPreCheck) (PandaLogProxy logger, const panda_seq_identifier *id, const panda_qual *forward, size_t forward_length,
                               const panda_qual *reverse, size_t reverse_length, void *user_data);
/* pandaseq-common.h:474-479 */
typedef bool (*PandaModuleCallback) (PandaAssembler assembler, PandaModule module, size_t rejected, void *data);
/* pandaseq-common.h:342-354 -- the shape panda_diff() takes for control/experiment */
typedef const panda_result_seq *(*PandaAssemble) (void *user_data, panda_seq_identifier *id,
                                                  const panda_qual *forward, size_t forward_length,
                                                  const panda_qual *reverse, size_t reverse_length);

/* pandaseq-common.h:594-605.  The two function pointers are HOST functions; the
 * device path never calls them.  It dispatches on class identity to the four
 * built-in device scorers and refuses any other class (PB_ERR_UNSUPPORTED). */
struct panda_algorithm_class {
	size_t data_size;
	const char *name;
	PandaAlgorithmCreate create;
	PandaDestroy data_destroy;
	PandaComputeOverlap overlap_probability;
	PandaComputeMatch match_probability;
	const double prob_unpaired;
};

#define PANDA_API 3
#define PANDA_DEFAULT_NUM_KMERS 2
/* misc.c:36-39; configure.ac:8 */
size_t panda_max_len(void);
#define PANDA_MAX_LEN (panda_max_len())

/* ======================================================================
 * Layer 1a: algorithms (pandaseq-algorithm.h:33-228, algo.c:27-133)
 * ====================================================================== */
extern PandaAlgorithmClass *panda_algorithms;           /* algo.c:85, sorted by name */
extern size_t panda_algorithms_length;                  /* algo.c:86 */
void panda_algorithm_register(PandaAlgorithmClass clazz);               /* algo.c:106-120 */
PandaAlgorithm panda_algorithm_new(PandaAlgorithmClass clazz);          /* algo.c:85-94 */
PandaAlgorithmClass panda_algorithm_class(PandaAlgorithm algo);         /* algo.c:39-42 */
void *panda_algorithm_data(PandaAlgorithm algo);                        /* algo.c:34-37 */
double panda_algorithm_quality_compare(PandaAlgorithm algorithm, const panda_qual *a, const panda_qual *b); /* algo.c:27-32 */
bool panda_algorithm_is_a(PandaAlgorithm algo, PandaAlgorithmClass clazz);  /* algo.c:44-48 */
PandaAlgorithm panda_algorithm_ref(PandaAlgorithm algo);                /* algo.c:50-60 */
void panda_algorithm_unref(PandaAlgorithm algo);                        /* algo.c:62-83 */

extern const struct panda_algorithm_class panda_algorithm_simple_bayes_class; /* algo_simple_bayes.c:100-108 */
PandaAlgorithm panda_algorithm_simple_bayes_new(void);                  /* algo_simple_bayes.c:110-115 */
double panda_algorithm_simple_bayes_get_error_estimation(PandaAlgorithm algorithm); /* algo_simple_bayes.c:117-124 */
void panda_algorithm_simple_bayes_set_error_estimation(PandaAlgorithm algorithm, double q); /* algo_simple_bayes.c:126-135 */

extern const struct panda_algorithm_class panda_algorithm_pear_class;   /* algo_pear.c:93-101 */
PandaAlgorithm panda_algorithm_pear_new(void);                          /* algo_pear.c:103-108 */
double panda_algorithm_pear_get_random_base_log_p(PandaAlgorithm algorithm); /* algo_pear.c:118-125 */
void panda_algorithm_pear_set_random_base_log_p(PandaAlgorithm algorithm, double log_p); /* algo_pear.c:110-116 */

extern const struct panda_algorithm_class panda_algorithm_rdp_mle_class; /* algo_rdp_mle.c:84-92 */
PandaAlgorithm panda_algorithm_rdp_mle_new(void);                        /* algo_rdp_mle.c:94-98 */

extern const struct panda_algorithm_class panda_algorithm_flash_class;   /* algo_flash.c:91-99 */
PandaAlgorithm panda_algorithm_flash_new(void);                          /* algo_flash.c:101-104 */

/* the three algorithms outside the north_star's four; same interface, count-based device scorers */
extern const struct panda_algorithm_class panda_algorithm_ea_util_class; /* algo_ea_util.c:79-87 */
PandaAlgorithm panda_algorithm_ea_util_new(void);                        /* algo_ea_util.c:89-92 */
extern const struct panda_algorithm_class panda_algorithm_stitch_class;  /* algo_stitch.c:79-87 */
PandaAlgorithm panda_algorithm_stitch_new(void);                         /* algo_stitch.c:89-92 */
extern const struct panda_algorithm_class panda_algorithm_uparse_class;  /* algo_uparse.c:100-108 */
PandaAlgorithm panda_algorithm_uparse_new(void);                         /* algo_uparse.c:110-115 */
double panda_algorithm_uparse_get_error_estimation(PandaAlgorithm algorithm); /* algo_uparse.c:117-124 */
void panda_algorithm_uparse_set_error_estimation(PandaAlgorithm algorithm, double q); /* algo_uparse.c:126-135 */

/* ======================================================================
 * Layer 1b: assembler (pandaseq-assembler.h:37-405, assembler.c, assembler_support.c)
 * ====================================================================== */
/* assembler_support.c:26-99.  `logger` may be NULL.  Returns NULL when no CUDA
 * device is usable or num_kmers != 2 (the reference's table indexing is only
 * self-consistent for 2). */
PandaAssembler panda_assembler_new(PandaNextSeq next, void *next_data, PandaDestroy next_destroy, PandaLogProxy logger);
PandaAssembler panda_assembler_new_kmer(PandaNextSeq next, void *next_data, PandaDestroy next_destroy, PandaLogProxy logger, size_t num_kmers);
PandaAssembler panda_assembler_ref(PandaAssembler assembler);           /* assembler_support.c:139-149 */
void panda_assembler_unref(PandaAssembler assembler);                   /* assembler_support.c:151-175 */
void panda_assembler_copy_configuration(PandaAssembler dest, PandaAssembler src); /* assembler_support.c:119-137 */

/* assembler.c:368-383.  A batch of one on the device; result is (transfer none),
 * valid until the next call on this assembler.  NULL = rejected, reason in the counters. */
const panda_result_seq *panda_assembler_assemble(PandaAssembler assembler, panda_seq_identifier *id,
                                                 const panda_qual *forward, size_t forward_length,
                                                 const panda_qual *reverse, size_t reverse_length);
/* assembler.c:350-366.  Pulls pairs from `next` in device-sized batches; returns
 * assembled pairs one at a time in input order, NULL when the source is dry. */
const panda_result_seq *panda_assembler_next(PandaAssembler assembler);

/* NEW (no reference counterpart; what a batching caller -- the replacement for
 * pool.c:71-108 do_assembly -- uses).  Assembles n pairs given as arrays of
 * pointers, calls `output` for every accepted pair in input order (may be NULL
 * to only update counters).  ids may be NULL.  Returns the number accepted, or
 * (size_t)-1 on a device/configuration error. */
size_t panda_assembler_assemble_batch(PandaAssembler assembler, size_t n, const panda_seq_identifier *ids,
                                      const panda_qual *const *forward, const size_t *forward_length,
                                      const panda_qual *const *reverse, const size_t *reverse_length,
                                      PandaOutputSeq output, void *output_data);

PandaAlgorithm panda_assembler_get_algorithm(PandaAssembler assembler);        /* assembler_support.c:177-180 */
/* Modules with host callbacks (pandaseq-module.h:47-57, 68-79, 111-113; pandaseq-assembler.h:106-124, 152-168; module.c:124-216).
 * Loading modules from shared objects (panda_module_load, libltdl) is out of scope; the built-in checks of the reference's
 * command line (-N, -l, -L) and its filter plugins run on the device through pb_config.filters instead. */
PandaModule panda_module_new(const char *name, PandaCheck check, PandaPreCheck precheck, void *user_data, PandaDestroy cleanup);
PandaModule panda_module_ref(PandaModule module);
void panda_module_unref(PandaModule module);
const char *panda_module_get_name(PandaModule module);
int panda_module_get_api(PandaModule module);                      /* PANDA_API (3) for constructed modules */
bool panda_assembler_add_module(PandaAssembler assembler, PandaModule module);
size_t panda_assembler_add_modules(PandaAssembler assembler, PandaModule *modules, size_t modules_length);
bool panda_assembler_foreach_module(PandaAssembler assembler, PandaModuleCallback callback, void *data);
void panda_assembler_module_stats(PandaAssembler assembler);       /* logging is out of scope: does nothing */
void panda_assembler_set_algorithm(PandaAssembler assembler, PandaAlgorithm algorithm); /* assembler_support.c:182-189 */
long panda_assembler_get_bad_read_count(PandaAssembler assembler);             /* assembler_support.c:191-194 */
long panda_assembler_get_count(PandaAssembler assembler);                      /* assembler_support.c:196-199 */
void panda_assembler_set_fail_alignment(PandaAssembler assembler, PandaFailAlign handler, void *handler_data, PandaDestroy handler_destroy); /* assembler_support.c:215-224 */
long panda_assembler_get_failed_alignment_count(PandaAssembler assembler);     /* assembler_support.c:226-229 */
panda_nt *panda_assembler_get_forward_primer(PandaAssembler assembler, size_t *length); /* assembler_support.c:231-237 */
void panda_assembler_set_forward_primer(PandaAssembler assembler, panda_nt *sequence, size_t length); /* assembler_support.c:201-213 */
size_t panda_assembler_get_forward_trim(PandaAssembler assembler);             /* assembler_support.c:239-242 */
void panda_assembler_set_forward_trim(PandaAssembler assembler, size_t trim);  /* assembler_support.c:244-249 */
size_t panda_assembler_get_longest_overlap(PandaAssembler assembler);          /* assembler_support.c:256-259 */
long panda_assembler_get_low_quality_count(PandaAssembler assembler);          /* assembler_support.c:266-269 */
int panda_assembler_get_minimum_overlap(PandaAssembler assembler);             /* assembler_support.c:271-274 */
void panda_assembler_set_minimum_overlap(PandaAssembler assembler, int overlap); /* assembler_support.c:276-282 */
int panda_assembler_get_maximum_overlap(PandaAssembler assembler);             /* assembler_support.c:284-287 */
void panda_assembler_set_maximum_overlap(PandaAssembler assembler, int overlap); /* assembler_support.c:289-295 */
const char *panda_assembler_get_name(PandaAssembler assembler);                /* assembler_support.c:297-302 */
void panda_assembler_set_name(PandaAssembler assembler, const char *name);     /* assembler_support.c:304-313 */
long panda_assembler_get_no_forward_primer_count(PandaAssembler assembler);    /* assembler_support.c:315-318 */
long panda_assembler_get_no_reverse_primer_count(PandaAssembler assembler);    /* assembler_support.c:320-323 */
size_t panda_assembler_get_num_kmer(PandaAssembler assembler);                 /* assembler_support.c:251-254 */
long panda_assembler_get_ok_count(PandaAssembler assembler);                   /* assembler_support.c:325-328 */
long panda_assembler_get_overlap_count(PandaAssembler assembler, size_t overlap); /* assembler_support.c:330-334 */
bool panda_assembler_get_primers_after(PandaAssembler assembler);              /* assembler_support.c:336-339 */
void panda_assembler_set_primers_after(PandaAssembler assembler, bool after);  /* assembler_support.c:341-345 */
panda_nt *panda_assembler_get_reverse_primer(PandaAssembler assembler, size_t *length); /* assembler_support.c:361-367 */
void panda_assembler_set_reverse_primer(PandaAssembler assembler, panda_nt *sequence, size_t length); /* assembler_support.c:347-359 */
size_t panda_assembler_get_reverse_trim(PandaAssembler assembler);             /* assembler_support.c:369-372 */
void panda_assembler_set_reverse_trim(PandaAssembler assembler, size_t trim);  /* assembler_support.c:374-379 */
long panda_assembler_get_slow_count(PandaAssembler assembler);                 /* assembler_support.c:381-384 */
double panda_assembler_get_threshold(PandaAssembler assembler);                /* assembler_support.c:386-389 */
void panda_assembler_set_threshold(PandaAssembler assembler, double threshold); /* assembler_support.c:391-397 */
PandaLogProxy panda_assembler_get_logger(PandaAssembler assembler);            /* assembler_support.c:261-264 */
double panda_assembler_get_primer_penalty(PandaAssembler assembler);           /* assembler_support.c:399-402 */
void panda_assembler_set_primer_penalty(PandaAssembler assembler, double threshold); /* assembler_support.c:404-410 */

/* pool.c:110-181 (pandaseq-args.h:155-169).  Drains the assembler's source through the device in batches and calls
 * output() for every assembled pair; consumes the assembler; threads/mux are accepted for signature compatibility. */
bool panda_run_pool(int threads, PandaAssembler assembler, PandaMux mux, PandaOutputSeq output, void *output_data, PandaDestroy output_destroy);

/* offset.c:103-112.  Runs the device primer scan on one read (batch of one). */
size_t panda_compute_offset_qual(double threshold, double penalty, bool reverse,
                                 const panda_qual *haystack, size_t haystack_length,
                                 const panda_nt *needle, size_t needle_length);

/* ======================================================================
 * Layer 2: batch / device layer
 * ====================================================================== */

typedef enum {
	PB_OK = 0,
	PB_ERR_NO_DEVICE = -1,    /* no CUDA device / driver: there is no CPU fallback */
	PB_ERR_CUDA = -2,         /* a CUDA call failed; see pb_last_error() */
	PB_ERR_UNSUPPORTED = -3,  /* configuration the device path does not implement */
	PB_ERR_ARGUMENT = -4,
	PB_ERR_NOMEM = -5
} pb_status;

#define PB_MAX_LEN 450  /* misc.c:38-41 */
#define PB_PHREDMAX 46

enum pb_algo { PB_SIMPLE_BAYES = 0, PB_PEAR = 1, PB_RDP_MLE = 2, PB_FLASH = 3, PB_EA_UTIL = 4, PB_STITCH = 5, PB_UPARSE = 6 };

/* Why a pair was not emitted, in assemble_seq order (assembler.c:252-348). */
enum pb_pair_status { PB_PAIR_OK = 0, PB_PAIR_BADR = 1, PB_PAIR_NOFP = 2, PB_PAIR_NORP = 3, PB_PAIR_NOALGN = 4, PB_PAIR_LOWQ = 5,
                      PB_PAIR_SKIP = 6 /* not a pair: a FASTQ record the reader drops (fastq.c:176) or one past the record that ended the stream; never counted */ };

/* One check on an assembled pair (see pb_config.filters). */
struct pb_filter {
	int32_t kind;         /* enum pb_filter_kind */
	int32_t ivalue;
	double dvalue;
	double dvalue2;       /* pear_test only: beta */
	double dvalue3;       /* pear_test only: cutoff */
};

/* Everything assemble_seq/align read from struct panda_assembler (assembler.h:28-79)
 * plus the algorithm's private data, as one plain struct.  panda_* objects are
 * flattened into this before every launch. */
typedef struct {
	int32_t algo;             /* enum pb_algo */
	int32_t post_primers;     /* primers_after (-a): locate the primers on the assembled sequence (assembler.c:300-333) */
	int64_t minoverlap;
	int64_t maxoverlap;
	int64_t num_kmers;        /* must be 2 */
	int64_t forward_trim;
	int64_t reverse_trim;
	int64_t forward_primer_length;
	int64_t reverse_primer_length;
	double threshold;         /* log space */
	double primer_penalty;
	double sb_q;              /* simple_bayes / uparse error estimation */
	double pear_random_base;  /* pear: log p of a random base */
	panda_nt forward_primer[PB_MAX_LEN];
	panda_nt reverse_primer[PB_MAX_LEN]; /* as the assembler stores it (already complemented) */
	/* Overhang trimmer in front of the assembler (hang.c:39-72, panda_trim_overhangs): the two sequences as that function
	 * receives them (args_hang.c:105-107: the reverse one already complemented); 0 length = off.  A pair whose overhang
	 * sequence is not found is dropped before it reaches the assembler (PB_PAIR_SKIP) unless hang_skip is set. */
	int64_t hang_forward_length;
	int64_t hang_reverse_length;
	int32_t hang_skip;
	int32_t nfilters;
	double hang_threshold;    /* passed to panda_compute_offset_qual as is (log space) */
	panda_nt hang_forward[PB_MAX_LEN];
	panda_nt hang_reverse[PB_MAX_LEN];
	/* Filters on the assembled pair, in module order (module.c:124-137): the first one that fails rejects the pair
	 * (status PB_PAIR_FILTERED + its index, counter PB_C_REJECTED + its index). */
	struct pb_filter filters[7];
} pb_config;
#define PB_MAX_FILTERS 7
enum pb_filter_kind {
	PB_FILTER_NONE = 0,
	PB_FILTER_NO_N = 1,            /* -N: degenerates == 0                          args_assembler.c:106-115 */
	PB_FILTER_SHORT = 2,           /* -l n: sequence_length >= n                    args_assembler.c:233-239 */
	PB_FILTER_LONG = 3,            /* -L n: sequence_length <= n                    args_assembler.c:268-275 */
	PB_FILTER_MIN_OVERLAPBITS = 4, /* min_overlapbits:bits  bits*ln2 <= estimated_overlap_probability, bits >= 0  plugin_min_overlapbits.c:17-50 */
	PB_FILTER_MISS_THE_POINT = 5,  /* completely_miss_the_point:n  overlap_mismatches <= n   plugin_completely_miss_the_point.c:9-16 */
	PB_FILTER_MIN_PHRED = 6,       /* min_phred:v  every base's panda_result_phred >= v      plugin_min_phred.c:8-22 */
	PB_FILTER_PEAR_TEST = 7        /* pear_test:alpha=,beta=,cutoff=  the statistical test of PEAR on (overlap, overlap_mismatches, read lengths):
	                                * dvalue = alpha, dvalue2 = beta, dvalue3 = cutoff in [0, 1]               plugin_pear_test.c:18-39 */
};
#define PB_PAIR_FILTERED 8

void pb_config_default(pb_config *cfg, int algo);

/* Counter vector; [PB_C_LONGEST] merges by max, everything else by sum
 * (the host-side STAT merge across shards). */
enum {
	PB_C_COUNT = 0, PB_C_OK, PB_C_LOWQ, PB_C_NOALGN, PB_C_BADR, PB_C_NOFP, PB_C_NORP, PB_C_SLOW,
	PB_C_LONGEST,
	PB_C_REJECTED = 9,     /* [9 .. 15]: pairs rejected by filter 0 .. 6 (the reference's per-module `rejected`, module.c:133) */
	PB_C_OVERLAPS = 16,
	PB_NCOUNTERS = 16 + 2 * PB_MAX_LEN
};
void pb_counters_merge(int64_t *dst, const int64_t *src);

/* --- packed device layout -------------------------------------------------
 * meta[i]   : {u32 off16; u16 flen; u16 rlen}, record i starts at reads + 16*off16
 * record    : [fwd nt, 4 bit/base, base 2k in the low nibble, padded to 4 B]
 *             [rev nt, same, stored in TEMPLATE order: element j is reverse[rlen-1-j]]
 *             [fwd qual, 1 B/base, raw char, padded to 4 B]
 *             [rev qual, template order, padded to 4 B]      then padded to 16 B
 */
typedef struct {
	uint32_t off16;
	uint16_t flen;
	uint16_t rlen;
} pb_pair_meta;

static inline size_t pb_record_bytes(size_t flen, size_t rlen) {
	size_t b = ((flen + 7) / 8) * 4 + ((rlen + 7) / 8) * 4 + ((flen + 3) / 4) * 4 + ((rlen + 3) / 4) * 4;
	return (b + 15) & ~(size_t) 15;
}

/* 32-byte result record, one per pair. */
typedef struct {
	uint8_t status;       /* enum pb_pair_status */
	uint8_t slow;         /* align() counted this pair as SLOW (assembler.c:135-137) */
	uint16_t overlap;
	uint16_t seq_len;
	uint16_t mismatches;
	uint16_t degenerates;
	uint16_t examined;
	uint16_t fwd_offset;
	uint16_t rev_offset;
	double quality;
	double est_prob;
} pb_pair_result;

typedef struct pb_context pb_context;

/* One context per (process, GPU).  Owns a stream, LUTs in HBM, staging buffers. */
pb_status pb_context_create(int device, pb_context **out);
void pb_context_destroy(pb_context *ctx);
const char *pb_last_error(void);
int pb_device_count(void);
/* the cudaStream_t the context launches on, as an opaque pointer (for event timing by the caller) */
void *pb_context_stream(pb_context *ctx);

/* Device-resident entry points: every pointer is a DEVICE pointer.
 * pb_pack_device:  flat AoS (panda_qual) -> packed records + meta.
 *   f_off/r_off have n+1 entries (element offsets into f_data/r_data, read order).
 *   rec_off16 has n entries (record offsets in 16 B units; a host/device exclusive
 *   scan of pb_record_bytes/16 -- pb_layout_host computes it). */
pb_status pb_pack_device(pb_context *ctx, size_t n,
                         const panda_qual *d_f_data, const uint64_t *d_f_off,
                         const panda_qual *d_r_data, const uint64_t *d_r_off,
                         const uint32_t *d_rec_off16, uint8_t *d_reads, pb_pair_meta *d_meta);

/* pb_assemble_device: the hot path.  seq_stride = row capacity in BASES (multiple of 16).
 * d_seq_nt [n][seq_stride/2] receives the merged read, 4 bit per base (panda_nt codes, base 2k
 * in the low nibble of byte k; may be NULL); d_seq_p [n][seq_stride] the per-base log p (may be
 * NULL: only FASTQ output and quality plugins need it); d_counters PB_NCOUNTERS
 * int64, accumulated.  max_read_len = longest read in the batch (selects the kernel's
 * shared-memory class; 0 = assume PB_MAX_LEN).  Asynchronous on the context's stream. */
pb_status pb_assemble_device(pb_context *ctx, const pb_config *cfg, size_t n, int max_read_len,
                             const uint8_t *d_reads, const pb_pair_meta *d_meta,
                             pb_pair_result *d_results, uint8_t *d_seq_nt, double *d_seq_p,
                             size_t seq_stride, int64_t *d_counters);
pb_status pb_synchronize(pb_context *ctx);
/* Diagnostics.  Configurations of the common kind (simple_bayesian / uparse / flash / pear, no primers/trims/trimmer, no
 * per-base log p) are assembled by a lane-per-pair kernel that hands the pairs it does not cover (a base that is not
 * A/C/G/T, a quality outside 0..46, no seed, ...) to the general warp-per-pair kernel.  This reports, since the context was
 * created, how many pairs were launched that way and how many of them were handed on.  Synchronises the device. */
pb_status pb_lanes_stats(pb_context *ctx, uint64_t *lanes_pairs, uint64_t *deferred_pairs);
/* 0: every configuration runs the general kernel; 1: the two-kernel path where it applies, its seeding by the diagonal sweep
 * (pb_sweep.cuh) where that applies; 2: the two-kernel path with the hash-join seeding kernel; -1 (default): as the environment
 * variables PANDASEQ_B200_LANES / PANDASEQ_B200_SWEEP say (unset = 1).  For A/B checks of the paths against each other. */
pb_status pb_set_lanes(pb_context *ctx, int mode);
/* Measurement.  With timing on, pb_assemble_device records CUDA events around its kernels on the context's stream;
 * pb_last_timing waits for the last call and returns their durations in milliseconds:
 *   kind 1: the general kernel alone, ms[2];
 *   kind 2: the two-kernel path -- ms[0] seeding (pb::seed_kernel), ms[1] lane-per-pair score + merge, ms[2] the general
 *           kernel over the pairs handed on;   kind 0: nothing recorded. */
pb_status pb_set_timing(pb_context *ctx, int on);
pb_status pb_last_timing(pb_context *ctx, int *kind, float ms[3]);

/* Host helpers for the layout (pure integer bookkeeping, no compute on reads). */
/* fills rec_off16[n] and returns total packed bytes */
size_t pb_layout_host(size_t n, const uint64_t *f_off, const uint64_t *r_off, uint32_t *rec_off16);

/* Host-buffer entry point (the e2e path): flat AoS batch in HOST memory, results in
 * HOST memory; H2D copy, pack, assemble and D2H copy happen inside, chunked and
 * double-buffered on the context's streams.  Output arrays may be NULL except results.
 * counters (host, PB_NCOUNTERS) is accumulated. */
pb_status pb_assemble_host(pb_context *ctx, const pb_config *cfg, size_t n,
                           const panda_qual *f_data, const uint64_t *f_off,
                           const panda_qual *r_data, const uint64_t *r_off,
                           pb_pair_result *results, uint8_t *seq_nt, double *seq_p,
                           size_t seq_stride, int64_t *counters);

/* pb_assemble_host with the per-base log p as 16-bit codes: seq_code[i][k] (row capacity seq_stride) is the index of base
 * k's log p in the table pb_posterior_table() fills -- every per-base log p the reference computes is an entry of one
 * 2 x 48 x 48 table (match_probability of the algorithm, qual_score[], qual_nn; assembler.c:162-243), so 2 bytes per base
 * cross the link instead of 8 and the caller expands them with the same doubles the device would have written.  Not
 * available together with primers-after or the min_phred filter (PB_ERR_ARGUMENT). */
#define PB_POSTERIOR_CODES (2 * 48 * 48)
pb_status pb_assemble_host_codes(pb_context *ctx, const pb_config *cfg, size_t n,
                                 const panda_qual *f_data, const uint64_t *f_off,
                                 const panda_qual *r_data, const uint64_t *r_off,
                                 pb_pair_result *results, uint8_t *seq_nt, uint16_t *seq_code,
                                 size_t seq_stride, int64_t *counters);
pb_status pb_posterior_table(const pb_config *cfg, double *table /* PB_POSTERIOR_CODES */);

/* The host path for callers that keep their reads in the packed layout above (a FASTQ parser can write it directly):
 * 458 instead of 620 bytes per 2x150 pair cross the link, and no pack kernel runs.  reads / meta as described under
 * "packed device layout", records in pair order and contiguous; max_read_len = longest read (0 = PB_MAX_LEN).
 * pb_pack_host is the layout done on the host (no arithmetic on the reads' content beyond `nt & 15`): reads must hold
 * pb_layout_host()'s total, meta n entries. */
pb_status pb_assemble_host_packed(pb_context *ctx, const pb_config *cfg, size_t n, int max_read_len,
                                  const uint8_t *reads, const pb_pair_meta *meta,
                                  pb_pair_result *results, uint8_t *seq_nt, size_t seq_stride, int64_t *counters);
void pb_pack_host(size_t n, const panda_qual *f_data, const uint64_t *f_off, const panda_qual *r_data, const uint64_t *r_off,
                  uint8_t *reads, pb_pair_meta *meta);
/* page-locked host memory (what the host path copies from / into without a staging copy) */
void *pb_host_alloc(size_t bytes);
void pb_host_free(void *p);

/* ======================================================================
 * Layer 2b: the stages either side of the hot path, on the device
 *   FASTQ text -> packed records   (fastq.c:44-193, linebuf.c:57-89, seqid.c:136-285, nt.c:48-124)
 *   assembled pairs -> FASTA/FASTQ text (output.c:85-126, nt.c:126-150, seqid.c:121-128)
 * ====================================================================== */
enum pb_tagging { PB_TAG_PRESENT = 0, PB_TAG_ABSENT = 1, PB_TAG_OPTIONAL = 2 };   /* PandaTagging, pandaseq-common.h:165-178 */
enum pb_idfmt { PB_IDFMT_UNKNOWN = 0, PB_IDFMT_SRA, PB_IDFMT_CASAVA_1_4, PB_IDFMT_CASAVA_1_7, PB_IDFMT_EBI_SRA, PB_IDFMT_CASAVA_CONVERTED }; /* PandaIdFmt, pandaseq-common.h:188-195 */
/* Why the reader stopped: the PandaCode fastq.c logs before it returns false for good. */
enum pb_fastq_error {
	PB_FQ_OK = 0,
	PB_FQ_ID_PARSE_FAILURE = 1, /* PANDA_CODE_ID_PARSE_FAILURE  fastq.c:126,133 */
	PB_FQ_NOT_PAIRED = 2,       /* PANDA_CODE_NOT_PAIRED        fastq.c:137 */
	PB_FQ_PREMATURE_EOF = 3,    /* PANDA_CODE_PREMATURE_EOF     fastq.c:58,69,84 */
	PB_FQ_BAD_NT = 4,           /* PANDA_CODE_BAD_NT            fastq.c:63 */
	PB_FQ_READ_TOO_LONG = 5,    /* PANDA_CODE_READ_TOO_LONG     fastq.c:75 */
	PB_FQ_PARSE_FAILURE = 6,    /* PANDA_CODE_PARSE_FAILURE     fastq.c:78 */
	PB_FQ_NO_QUALITY_INFO = 7,  /* PANDA_CODE_NO_QUALITY_INFO   fastq.c:95 */
	PB_FQ_LINE_TOO_LONG = 8     /* a line of 4500 bytes or more: linebuf.c:25,65 returns NULL, the stream just ends */
};
#define PB_FQ_LINE_MAX 4500

/* A parsed identifier, 48 bytes: the strings are substrings of the forward header line. */
typedef struct {
	uint32_t hdr_off;     /* offset, in the forward text handed to the parser, of the character after '@' */
	uint16_t hdr_len;     /* length of the header line after that character (CR stripped) */
	uint8_t fmt;          /* enum pb_idfmt */
	uint8_t reserved;
	uint16_t inst_off, inst_len;   /* relative to hdr_off; for the SRA formats the instrument is "%cRR%d" of `sra` */
	uint16_t run_off, run_len;
	uint16_t fc_off, fc_len;
	uint16_t tag_off, tag_len;
	int32_t lane, tile, x, y;
	int32_t sra;
	int32_t mate;         /* what panda_seqid_parse returns (the read direction) */
} pb_seq_id;
/* Expand into the reference's 368-byte struct (host helper; `fwd_text` is the text the parser saw). */
void pb_seq_id_expand(const pb_seq_id *id, const char *fwd_text, panda_seq_identifier *out);

typedef struct {
	uint64_t records;      /* complete 4-line records present in BOTH texts */
	uint64_t limit;        /* records before the one that ended the stream (== records when none did) */
	uint64_t pairs;        /* pairs delivered: limit minus the records with an empty forward read (fastq.c:176) */
	uint64_t consumed_fwd; /* bytes of each text that belong to records [0, records) */
	uint64_t consumed_rev;
	int32_t error;         /* enum pb_fastq_error of record `limit`, PB_FQ_OK when the data simply ran out */
	int32_t max_read_len;  /* longest read among the records (selects the assemble kernel's class) */
	uint32_t stride16;     /* packed records are laid out at a fixed stride of 16*stride16 bytes */
	uint32_t reserved;
} pb_fastq_info;

/* pb_fastq_parse_device: every pointer is a DEVICE pointer, the call returns after the (small) pb_fastq_info came
 * back.  d_fwd/d_rev: FASTQ text (16-byte aligned).  Record i of the result is FASTQ record i: d_meta[i].flen ==
 * 0xFFFF marks a record that is not a pair (pb_assemble_device gives it PB_PAIR_SKIP and does not count it); assemble
 * records [0, info.limit).  d_reads must hold records * pb_record_bytes(PB_MAX_LEN, PB_MAX_LEN) bytes unless the caller knows better
 * (info.stride16 tells what was used), d_meta / d_ids `max_records` entries.  Text that does not end on a record
 * boundary is the caller's to resubmit (consumed_*). */
pb_status pb_fastq_parse_device(pb_context *ctx, const char *d_fwd, size_t fwd_bytes, const char *d_rev, size_t rev_bytes,
                                int qualmin, int policy, size_t max_records,
                                uint8_t *d_reads, size_t reads_capacity, pb_pair_meta *d_meta, pb_seq_id *d_ids, pb_fastq_info *info);

enum pb_out_format { PB_OUT_FASTA = 0, PB_OUT_FASTQ = 1 };   /* panda_output_fasta / panda_output_fastq */
/* pb_format_device: text of every PB_PAIR_OK record, in record order, written to d_text (capacity bytes); FASTQ needs
 * d_seq_p.  d_fwd is the forward text the ids point into.  *text_bytes receives the length (a larger value than
 * capacity means nothing was written: call again with enough room). */
pb_status pb_format_device(pb_context *ctx, int format, size_t n, const pb_pair_result *d_results, const uint8_t *d_seq_nt,
                           const double *d_seq_p, size_t seq_stride, const pb_seq_id *d_ids, const char *d_fwd,
                           char *d_text, size_t capacity, size_t *text_bytes);

/* The whole chain with HOST buffers: two FASTQ texts in, assembled FASTA/FASTQ text out; H2D copies, parse, assemble,
 * format and the D2H copy happen inside, in chunks on two streams.  `final` != 0: the texts are complete files (a
 * trailing partial record is an error); == 0: the caller resubmits what consumed_* leaves.  out_text may be NULL to
 * only count.  counters (PB_NCOUNTERS, accumulated) as pb_assemble_host. */
typedef struct {
	uint64_t records, pairs, consumed_fwd, consumed_rev, out_bytes;
	int32_t error;         /* enum pb_fastq_error */
	int32_t reserved;
} pb_stream_info;
pb_status pb_fastq_assemble_host(pb_context *ctx, const pb_config *cfg, int qualmin, int policy, int out_format,
                                 const char *fwd, size_t fwd_bytes, const char *rev, size_t rev_bytes, int final,
                                 char *out_text, size_t out_capacity, int64_t *counters, pb_stream_info *info);

/* The LUTs the kernels use (regenerated with the reference's formulas and "%g"
 * rounding, mktable.c:23-155 + tablebuilder.c:73-183), exposed for the table parity test. */
typedef struct {
	double qual_nn;
	double match_sb[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double mismatch_sb[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double match_pear[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double mismatch_pear[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double mismatch_rdp[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double mismatch_rdp_asm[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double match_uparse[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double mismatch_uparse[PB_PHREDMAX + 1][PB_PHREDMAX + 1];
	double score[PB_PHREDMAX + 1];
	double score_err[PB_PHREDMAX + 1];
} pb_tables;
const pb_tables *pb_get_tables(void);

#ifdef __cplusplus
}
#endif
#endif
