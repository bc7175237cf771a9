/* api_bench.c -- read pairs per second through the reference's own entry points: a PandaNextSeq source, a PandaAssembler,
 * panda_run_pool() and a PandaOutputSeq callback (pool.c:71-181, assembler.c:350-366).
 *
 * ONE source file for both sides of the comparison.  Built against include/pandaseq_b200.h and linked with
 * libpandaseq_b200.so it measures this library (pandaseq_b200/csrc/Makefile -> pandaseq_b200/api_bench); built with
 * -DAPI_BENCH_REFERENCE against the reference's pandaseq.h and linked with the compiled reference it measures the reference
 * (oracle/Makefile -> oracle/_ref/api_bench_ref; the reference needs a PandaMux to share a source between threads, this
 * library shares the assembler's own source).  bench.py --path api runs both.
 *
 *   api_bench <pairs> <read length> <threads> [check]
 *
 * Synthetic 2 x L pairs of a random template (insert L+30 .. 2L-20), declining qualities, substitution errors at the rate the
 * quality states: the shape of BASELINE config 2, generated here so that the program needs no input file.  The output callback
 * does what a writer would have to: it reads every base and log p of the result, into per-thread sums.  Prints one JSON line. */
#define _POSIX_C_SOURCE 200809L
#ifdef API_BENCH_REFERENCE
#include <pandaseq.h>
#include <pandaseq-mux.h>
#else
#include <pandaseq_b200.h>
#endif
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct source {
	size_t n, i, len;
	panda_qual *f, *r;		/* n x len each */
	pthread_mutex_t lock;
};

static bool next_pair(panda_seq_identifier *id, const panda_qual **f, size_t *fl, const panda_qual **r, size_t *rl, void *user) {
	struct source *s = user;
	if (s->i >= s->n)
		return false;
	memset(id, 0, sizeof *id);
	id->x = (int) (s->i & 0x7FFFFFFF);
	*f = s->f + s->i * s->len;
	*r = s->r + s->i * s->len;
	*fl = *rl = s->len;
	s->i++;
	return true;
}

/* What the output callback adds up, per calling thread (a slot per thread, summed at the end): a writer that took one lock per
 * result would measure the lock, not the library -- the reference's own writer keeps a buffer per thread for that reason
 * (writer.c:93, a pthread key). */
struct sink_slot {
	unsigned long long pairs, bases;
	double psum;
	char pad[40];
};
struct sink {
	struct sink_slot slot[256];
};
static int next_slot;
static __thread int my_slot = -1;

static bool count_output(const panda_result_seq *seq, void *user) {
	struct sink *k = user;
	unsigned long long bases = 0;
	double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
	const panda_result *r = seq->sequence;
	const size_t n = seq->sequence_length;
	size_t i = 0;
	if (my_slot < 0)
		my_slot = __atomic_fetch_add(&next_slot, 1, __ATOMIC_RELAXED) & 255;
	/* every base and every log p is read, as a FASTQ writer would; four partial sums so that the additions overlap */
	for (; i + 4 <= n; i += 4) {
		bases += ((r[i].nt & 15) != 0) + ((r[i + 1].nt & 15) != 0) + ((r[i + 2].nt & 15) != 0) + ((r[i + 3].nt & 15) != 0);
		p0 += r[i].p;
		p1 += r[i + 1].p;
		p2 += r[i + 2].p;
		p3 += r[i + 3].p;
	}
	for (; i < n; i++) {
		bases += (r[i].nt & 15) != 0;
		p0 += r[i].p;
	}
	k->slot[my_slot].pairs++;
	k->slot[my_slot].bases += bases;
	k->slot[my_slot].psum += (p0 + p1) + (p2 + p3);
	return true;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd(void) {
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return (uint32_t) (rng_state >> 32);
}

static void generate(struct source *s) {
	static const char comp[9] = { 0, 8, 4, 0, 2, 0, 0, 0, 1 };
	const size_t L = s->len;
	char *tmpl = malloc(2 * L);
	double perr[64];
	for (int q = 0; q < 64; q++)
		perr[q] = pow(10.0, -q / 10.0);
	for (size_t p = 0; p < s->n; p++) {
		const size_t T = L + 30 + rnd() % (L - 50 + 1);
		for (size_t k = 0; k < T; k++)
			tmpl[k] = (char) (1 << (rnd() & 3));
		for (int side = 0; side < 2; side++) {
			panda_qual *out = (side ? s->r : s->f) + p * L;
			for (size_t k = 0; k < L; k++) {
				int q = (int) (36.0 - 12.0 * (double) k / (double) L + ((int) (rnd() % 9) - 4));
				q = q < 2 ? 2 : (q > 41 ? 41 : q);
				/* forward read: template from its start; reverse read: the other strand from the template's end, stored
				 * complemented in read order (fastq.c:154), i.e. the template's bases from the end backwards */
				char nt = side ? tmpl[T - 1 - k] : tmpl[k];
				if ((rnd() & 0xFFFFFF) < perr[q] * 16777216.0) {
					int b = 0;
					while ((1 << b) != nt)
						b++;
					nt = (char) (1 << ((b + 1 + rnd() % 3) & 3));
				}
				(void) comp;
				out[k].nt = nt;
				out[k].qual = (char) q;
			}
		}
	}
	free(tmpl);
}

static double now(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

int main(int argc, char **argv) {
	struct source src;
	static struct sink sink;
	PandaAssembler a, keep;
	int threads;
	double t0, t1;
	if (argc < 4) {
		fprintf(stderr, "usage: %s pairs read_length threads\n", argv[0]);
		return 2;
	}
	memset(&src, 0, sizeof src);
	src.n = (size_t) atol(argv[1]);
	src.len = (size_t) atol(argv[2]);
	threads = atoi(argv[3]);
	if (src.len < 60 || src.len > 400 || src.n == 0 || threads < 1)
		return 2;
	src.f = malloc(src.n * src.len * sizeof(panda_qual));
	src.r = malloc(src.n * src.len * sizeof(panda_qual));
	if (src.f == NULL || src.r == NULL)
		return 2;
	generate(&src);
#ifdef API_BENCH_REFERENCE
	{
		PandaWriter w = panda_writer_new_null();
		PandaLogProxy logger = panda_log_proxy_new(w);
		PandaMux mux = NULL;
		panda_writer_unref(w);
		panda_debug_flags = 0;
		if (threads > 1) {
			mux = panda_mux_new(next_pair, &src, NULL, logger);
			a = panda_mux_create_assembler(mux);
		} else {
			a = panda_assembler_new(next_pair, &src, NULL, logger);
		}
		if (a == NULL)
			return 3;
		keep = panda_assembler_ref(a);
		t0 = now();
		panda_run_pool(threads, a, mux, count_output, &sink, NULL);
		t1 = now();
		panda_log_proxy_unref(logger);
	}
#else
	a = panda_assembler_new(next_pair, &src, NULL, NULL);
	if (a == NULL) {
		fprintf(stderr, "no assembler: %s\n", pb_last_error());
		return 3;
	}
	keep = panda_assembler_ref(a);
	{	/* a first small run so that the contexts, the pinned staging and the kernels' first launch are not in the timing */
		struct source warm = src;
		PandaAssembler w = panda_assembler_new(next_pair, &warm, NULL, NULL);
		static struct sink ws;
		warm.n = src.n < 400000 ? src.n : 400000;
		panda_run_pool(threads, w, NULL, count_output, &ws, NULL);
	}
	t0 = now();
	panda_run_pool(threads, a, NULL, count_output, &sink, NULL);
	t1 = now();
#endif
	unsigned long long pairs = 0, bases = 0;
	double psum = 0;
	for (int k = 0; k < 256; k++) {
		pairs += sink.slot[k].pairs;
		bases += sink.slot[k].bases;
		psum += sink.slot[k].psum;
	}
	printf("{\"pairs\": %zu, \"read_length\": %zu, \"threads\": %d, \"seconds\": %.6f, \"mpairs_per_s\": %.4f, \"ok\": %llu, "
	       "\"count\": %ld, \"bases\": %llu, \"psum\": %.6f}\n",
	       src.n, src.len, threads, t1 - t0, (double) src.n / (t1 - t0) / 1e6, pairs,
	       panda_assembler_get_count(keep), bases, psum);
	panda_assembler_unref(keep);
	return 0;
}
