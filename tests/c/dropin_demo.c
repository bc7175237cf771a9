/* dropin_demo.c -- a plain C program written against the reference's API shape, compiled against
 * include/pandaseq_b200.h and linked with -lpandaseq_b200.  It is the loop of pool.c:71-108 (pull pairs from a
 * PandaNextSeq source, assemble, hand each result to an output callback, print the STAT block) with the GPU behind
 * panda_assembler_next().  Usage: dropin_demo <pairs.bin> <algo> [threads]; pairs.bin is the flat dump written by the test.
 * With threads > 1 panda_run_pool fans the work out over that many workers (one GPU each while there are GPUs), which call
 * the output function concurrently, as the reference's pool does: the callback below takes a lock. */
#include <pandaseq_b200.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct source {
	size_t n, i;
	uint64_t *f_off, *r_off;
	panda_qual *f_data, *r_data;
};

static bool next_pair(panda_seq_identifier *id, const panda_qual **f, size_t *fl, const panda_qual **r, size_t *rl, void *user) {
	struct source *s = user;
	if (s->i >= s->n)
		return false;
	memset(id, 0, sizeof *id);
	id->x = (int) s->i;
	*f = s->f_data + s->f_off[s->i];
	*fl = (size_t) (s->f_off[s->i + 1] - s->f_off[s->i]);
	*r = s->r_data + s->r_off[s->i];
	*rl = (size_t) (s->r_off[s->i + 1] - s->r_off[s->i]);
	s->i++;
	return true;
}

static pthread_mutex_t out_lock = PTHREAD_MUTEX_INITIALIZER;

static bool print_fasta(const panda_result_seq *seq, void *user) {
	static const char letters[] = "NACMGRSVTWYHKDBN";
	FILE *out = user;
	pthread_mutex_lock(&out_lock);
	fprintf(out, ">pair%d;overlap=%zu;mismatches=%zu;q=%.6f\n", seq->name.x, seq->overlap, seq->overlap_mismatches, seq->quality);
	for (size_t k = 0; k < seq->sequence_length; k++)
		fputc(letters[seq->sequence[k].nt & 15], out);
	fputc('\n', out);
	pthread_mutex_unlock(&out_lock);
	return true;
}

static void *slurp(FILE *f, size_t bytes) {
	void *p = malloc(bytes ? bytes : 1);
	if (p == NULL || fread(p, 1, bytes, f) != bytes) {
		fprintf(stderr, "short read\n");
		exit(2);
	}
	return p;
}

int main(int argc, char **argv) {
	struct source src;
	uint64_t hdr[3];
	PandaAssembler a, b;
	PandaAlgorithm algo = NULL;
	FILE *f;
	if (argc < 3 || (f = fopen(argv[1], "rb")) == NULL) {
		fprintf(stderr, "usage: %s pairs.bin simple_bayesian|pear|rdp_mle|flash\n", argv[0]);
		return 2;
	}
	if (fread(hdr, sizeof hdr, 1, f) != 1)
		return 2;
	src.n = hdr[0];
	src.i = 0;
	src.f_off = slurp(f, (src.n + 1) * 8);
	src.r_off = slurp(f, (src.n + 1) * 8);
	src.f_data = slurp(f, hdr[1] * sizeof(panda_qual));
	src.r_data = slurp(f, hdr[2] * sizeof(panda_qual));
	fclose(f);
	for (size_t k = 0; k < panda_algorithms_length; k++)
		if (strcmp(panda_algorithms[k]->name, argv[2]) == 0)
			algo = panda_algorithms[k]->create(NULL);
	if (algo == NULL) {
		fprintf(stderr, "unknown algorithm %s\n", argv[2]);
		return 2;
	}
	a = panda_assembler_new(next_pair, &src, NULL, NULL);
	if (a == NULL) {
		fprintf(stderr, "no assembler: %s\n", pb_last_error());
		return 3;               /* no GPU: there is no CPU fallback */
	}
	panda_assembler_set_algorithm(a, algo);
	panda_algorithm_unref(algo);
	panda_assembler_set_threshold(a, 0.6);
	panda_assembler_set_minimum_overlap(a, 2);
	b = panda_assembler_ref(a);     /* keep the counters readable after run_pool consumed its reference */
	bool any = panda_run_pool(argc > 3 ? atoi(argv[3]) : 1, a, NULL, print_fasta, stdout, NULL);
	fprintf(stderr, "STAT\tREADS\t%ld\nSTAT\tNOALGN\t%ld\nSTAT\tLOWQ\t%ld\nSTAT\tBADR\t%ld\nSTAT\tSLOW\t%ld\nSTAT\tOK\t%ld\nSTAT\tLONGEST\t%zu\n",
	        panda_assembler_get_count(b), panda_assembler_get_failed_alignment_count(b), panda_assembler_get_low_quality_count(b),
	        panda_assembler_get_bad_read_count(b), panda_assembler_get_slow_count(b), panda_assembler_get_ok_count(b),
	        panda_assembler_get_longest_overlap(b));
	panda_assembler_unref(b);
	return any ? 0 : 1;
}
