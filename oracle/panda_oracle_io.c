/* panda_oracle_io.c -- CPU restatement of the two stages either side of the assembly hot path:
 * FASTQ text -> panda_qual pairs (fastq.c, linebuf.c, seqid.c) and assembled pair -> FASTA/FASTQ
 * text (output.c, nt.c:126-150, seqid.c:121-128).
 *
 * TEST INFRASTRUCTURE ONLY (see panda_oracle.h).  Pinned against the compiled reference by
 * tests/test_io_oracle_vs_ref.py and tests/golden/io_*.npz.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "panda_oracle.h"

/* ---- nt.c:25-105: ASCII <-> 4-bit codes -------------------------------------------------- */
/* index = c & 0x1F ('@'..'_' and their lowercase twins); 0 = not a nucleotide letter */
static const unsigned char fwd_code[32] = {
	0, 1, 14, 2, 13, 0, 0, 4, 11, 0, 0, 12, 0, 3, 15, 0,
	0, 0, 5, 6, 8, 8, 7, 9, 15, 10, 0, 0, 0, 0, 0, 0
};

static unsigned char complement4(unsigned char c) {
	/* A<->T, C<->G bitwise: reverse the four bits */
	return (unsigned char) (((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3));
}

unsigned char po_nt_from_ascii(char c, int complement) {
	unsigned char v = fwd_code[(int) c & 0x1F];
	return complement ? complement4(v) : v;
}

char po_nt_to_ascii(char nt) {
	static const char letters[16] = { 'N', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N' };
	if (nt < 0 || nt > 15)
		return 'N';
	return letters[(int) nt];
}

/* ---- seqid.c:136-285 ------------------------------------------------------------------------- */
static int is_delim(char c) {
	return c == '\0' || c == ':' || c == '#' || c == '/' || c == ' ';
}

typedef struct {
	const char *p;
} cursor;

/* a "chunk": characters up to the next delimiter.  must_have: fail when already at the end of the string */
static int take_str(cursor *c, char *target, size_t cap, int must_have) {
	size_t n = 0;
	if (must_have && *c->p == '\0')
		return 0;
	while (!is_delim(*c->p)) {
		if (n >= cap)                     /* seqid.c:153 tests `>`: the reference accepts one more byte than the field
			                                   * holds and writes it into the next member; that case is refused here */
			return 0;
		target[n++] = *c->p++;
	}
	target[n] = '\0';
	return 1;
}

static int take_int(cursor *c, int *value) {
	unsigned v = 0;
	if (*c->p == '\0')
		return 0;
	while (!is_delim(*c->p)) {
		if (*c->p < '0' || *c->p > '9')
			return 0;
		v = 10u * v + (unsigned) (*c->p - '0');     /* wraps like the reference's int does on x86 */
		c->p++;
	}
	*value = (int) v;
	return 1;
}

static int take_sra_int(cursor *c, int *value) {
	unsigned v = 0;
	while (*c->p != '\0' && *c->p != '.' && *c->p != ' ') {
		if (*c->p < '0' || *c->p > '9')
			return 0;
		v = 10u * v + (unsigned) (*c->p - '0');
		c->p++;
	}
	*value = (int) v;
	return 1;
}

static int skip_chunk(cursor *c) {
	if (*c->p == '\0')
		return 0;
	while (!is_delim(*c->p))
		c->p++;
	return 1;
}

static int push(cursor *c) {
	if (*c->p == '\0')
		return 0;
	c->p++;
	return 1;
}

static int take_tag(cursor *c, char *tag) {
	size_t n = 0;
	tag[0] = '\0';
	while (!is_delim(*c->p)) {
		if (n >= PO_TAG_LEN)              /* seqid.c:222,267 */
			return 0;
		tag[n++] = *c->p++;
		if (n < PO_TAG_LEN)
			tag[n] = '\0';
	}
	return 1;
}

static int tag_policy_ok(const po_seq_identifier *id, int policy) {
	if (policy == PO_TAG_OPTIONAL)
		return 1;
	return policy == ((id->tag[0] == '\0') ? PO_TAG_ABSENT : PO_TAG_PRESENT);
}

int po_seqid_parse(po_seq_identifier *id, const char *input, int policy, int *format) {
	cursor c = { input };
	int v, fmt_dummy;
	if (format == NULL)
		format = &fmt_dummy;
	if (strlen(input) > 3 && (input[0] == 'E' || input[0] == 'S') && input[1] == 'R' && input[2] == 'R') {
		*format = input[0] == 'S' ? PO_IDFMT_SRA : PO_IDFMT_EBI_SRA;
		c.p += 3;
		memset(id, 0, sizeof *id);
		if (!take_sra_int(&c, &v) || !push(&c))
			return 0;
		sprintf(id->instrument, "%cRR%d", (int) input[0], v);
		if (!take_sra_int(&c, &v) || !push(&c))
			return 0;
		id->lane = v;
		if (!push(&c))
			return 0;
		return 1;
	}
	if (strchr(input, '/') != NULL) {
		size_t colons = 0;
		char before;
		for (const char *s = input; *s != '\0' && *s != '#'; s++)
			if (*s == ':')
				colons++;
		if (colons == 6) {
			*format = PO_IDFMT_CASAVA_CONVERTED;
			if (!take_str(&c, id->instrument, sizeof id->instrument, 1) || !push(&c)) return 0;
			if (!take_str(&c, id->run, sizeof id->run, 1) || !push(&c)) return 0;
			if (!take_str(&c, id->flowcell, sizeof id->flowcell, 1) || !push(&c)) return 0;
		} else {
			*format = PO_IDFMT_CASAVA_1_4;
			id->run[0] = '\0';
			id->flowcell[0] = '\0';
			if (!take_str(&c, id->instrument, sizeof id->instrument, 1) || !push(&c)) return 0;
		}
		if (!take_int(&c, &id->lane) || !push(&c)) return 0;
		if (!take_int(&c, &id->tile) || !push(&c)) return 0;
		if (!take_int(&c, &id->x) || !push(&c)) return 0;
		if (!take_int(&c, &id->y) || !push(&c)) return 0;
		before = c.p[-1];
		id->tag[0] = '\0';
		if (before == '#') {
			if (!take_tag(&c, id->tag) || !push(&c))
				return 0;
		}
		if (!tag_policy_ok(id, policy))
			return 0;
		if (!take_int(&c, &v))
			return 0;
		return v;
	}
	{
		int mate;
		*format = PO_IDFMT_CASAVA_1_7;
		if (!take_str(&c, id->instrument, sizeof id->instrument, 1) || !push(&c)) return 0;
		if (!take_str(&c, id->run, sizeof id->run, 1) || !push(&c)) return 0;
		if (!take_str(&c, id->flowcell, sizeof id->flowcell, 1) || !push(&c)) return 0;
		if (!take_int(&c, &id->lane) || !push(&c)) return 0;
		if (!take_int(&c, &id->tile) || !push(&c)) return 0;
		if (!take_int(&c, &id->x) || !push(&c)) return 0;
		if (!take_int(&c, &id->y) || !push(&c)) return 0;
		if (!take_int(&c, &mate) || !push(&c)) return 0;
		if (!skip_chunk(&c) || !push(&c)) return 0;                             /* "filtered" flag: any chunk */
		if (!take_int(&c, &v) || !push(&c)) return 0;                           /* control bits */
		if (!take_tag(&c, id->tag))
			return 0;
		if (!tag_policy_ok(id, policy))
			return 0;
		return mate;
	}
}

static int seqid_equal(const po_seq_identifier *a, const po_seq_identifier *b) {
	return a->lane == b->lane && a->tile == b->tile && a->x == b->x && a->y == b->y
		&& strncmp(a->instrument, b->instrument, sizeof a->instrument) == 0 && strncmp(a->run, b->run, sizeof a->run) == 0
		&& strncmp(a->flowcell, b->flowcell, sizeof a->flowcell) == 0 && strncmp(a->tag, b->tag, sizeof a->tag) == 0;
}

/* ---- linebuf.c:57-89 over an in-memory buffer ---------------------------------------------------
 * Returns 1 and a NUL-terminated copy of the next line in `line` (the reference hands out a C string, so a
 * NUL inside the line ends it early), 0 at the end of the data.  A final line that lacks its '\n' is not
 * returned (the reference then reads one byte past its data; with zeroed memory that is the same answer),
 * nor is a line that does not end inside the 4500-byte window. */
typedef struct {
	const char *data;
	size_t len, pos;
} mem_lines;

static int next_line(mem_lines *m, char *line) {
	size_t avail = m->len - m->pos, n;
	const char *nl;
	if (avail == 0)
		return 0;
	nl = memchr(m->data + m->pos, '\n', avail < PO_FQ_LINE_MAX ? avail : PO_FQ_LINE_MAX);
	if (nl == NULL)
		return 0;
	n = (size_t) (nl - (m->data + m->pos));
	memcpy(line, m->data + m->pos, n);
	if (n > 0 && line[n - 1] == '\r')
		n--;
	line[n] = '\0';
	m->pos += (size_t) (nl - (m->data + m->pos)) + 1;
	return 1;
}

/* fastq.c:44-102 */
static int read_seq(mem_lines *m, char *line, po_qual *buffer, int complement, int qualmin, size_t *length) {
	size_t pos = 0, qpos = 0;
	const char *in;
	if (!next_line(m, line))
		return PO_FQ_PREMATURE_EOF;
	for (in = line; *in != '\0' && pos < PO_MAX_LEN; in++) {
		if ((buffer[pos++].nt = (char) po_nt_from_ascii(*in, complement)) == '\0')
			return PO_FQ_BAD_NT;
	}
	if (!next_line(m, line))
		return PO_FQ_PREMATURE_EOF;
	if (line[0] != '+')
		return po_nt_from_ascii(line[0], complement) != 0 ? PO_FQ_READ_TOO_LONG : PO_FQ_PARSE_FAILURE;
	if (!next_line(m, line))
		return PO_FQ_PREMATURE_EOF;
	for (in = line; *in != '\0'; in++) {
		int val = (int) *in, q;
		if (val < qualmin)
			q = 0;
		else
			q = (val > qualmin + PO_PHREDMAX ? PO_PHREDMAX : val) - qualmin;     /* fastq.c:44: the clamp is applied before the offset */
		if (qpos < PO_MAX_LEN)
			buffer[qpos].qual = (char) q;
		qpos++;
	}
	if (qpos != pos)
		return PO_FQ_NO_QUALITY_INFO;
	*length = pos;
	return PO_FQ_OK;
}

/* fastq.c:106-193 without the index-read file */
int po_fastq_parse(const char *fwd, size_t fwd_len, const char *rev, size_t rev_len, int qualmin, int policy,
                   size_t max_pairs, po_fastq_out *out) {
	mem_lines mf = { fwd, fwd_len, 0 }, mr = { rev, rev_len, 0 };
	char *line = malloc(PO_FQ_LINE_MAX + 2);
	po_qual fbuf[PO_MAX_LEN], rbuf[PO_MAX_LEN];
	size_t n = 0, records = 0;
	uint64_t fo = 0, ro = 0;
	out->error = PO_FQ_OK;
	if (out->f_off) out->f_off[0] = 0;
	if (out->r_off) out->r_off[0] = 0;
	while (n < max_pairs) {
		po_seq_identifier id, rid;
		int fmt, fdir, rdir, rc;
		size_t flen = 0, rlen = 0;
		memset(&id, 0, sizeof id);
		memset(&rid, 0, sizeof rid);
		if (!next_line(&mf, line))
			break;
		if (line[0] == '\0' || (fdir = po_seqid_parse(&id, line + 1, policy, &fmt)) == 0) {
			out->error = PO_FQ_ID_PARSE_FAILURE;
			break;
		}
		if (!next_line(&mr, line))
			break;
		if (line[0] == '\0' || (rdir = po_seqid_parse(&rid, line + 1, policy, NULL)) == 0) {
			out->error = PO_FQ_ID_PARSE_FAILURE;
			break;
		}
		if (!seqid_equal(&id, &rid) || (fmt != PO_IDFMT_SRA && fmt != PO_IDFMT_EBI_SRA && rdir == fdir)) {
			out->error = PO_FQ_NOT_PAIRED;
			break;
		}
		if ((rc = read_seq(&mf, line, fbuf, 0, qualmin, &flen)) != PO_FQ_OK || (rc = read_seq(&mr, line, rbuf, 1, qualmin, &rlen)) != PO_FQ_OK) {
			out->error = rc;
			break;
		}
		records++;
		if (flen == 0)
			continue;                   /* fastq.c:176: an empty forward read is skipped, not reported */
		if (out->ids) out->ids[n] = id;
		if (out->f_data) memcpy(out->f_data + fo, fbuf, flen * sizeof(po_qual));
		if (out->r_data) memcpy(out->r_data + ro, rbuf, rlen * sizeof(po_qual));
		fo += flen;
		ro += rlen;
		n++;
		if (out->f_off) out->f_off[n] = fo;
		if (out->r_off) out->r_off[n] = ro;
	}
	out->n = n;
	out->records = records;
	free(line);
	return 0;
}

/* ---- nt.c:126-150 ---------------------------------------------------------------------------- */
char po_result_phred(double p) {
	const po_tables *t = po_get_tables();
	char lower = 0, upper = PO_PHREDMAX;
	if (p <= t->score[0])
		return 1;
	while (lower < upper) {
		char mid = (char) (lower + (upper - lower) / 2);
		if (t->score[(int) mid] == p)
			return mid;
		if (mid == lower)
			return lower;
		if (t->score[(int) mid] > p)
			upper = mid;
		else if (t->score[(int) mid] < p)
			lower = (char) (mid + 1);
	}
	return lower;
}

/* ---- output.c:85-126 + seqid.c:121-128 ------------------------------------------------------------
 * Appends the FASTA (fastq == 0) or FASTQ record of one assembled pair to dst; returns the bytes written. */
size_t po_format_record(char *dst, int fastq, const po_seq_identifier *id, double quality,
                        const uint8_t *seq_nt, const double *seq_p, size_t seq_len) {
	char *o = dst;
	if (seq_len == 0)
		return 0;
	*o++ = fastq ? '@' : '>';
	o += sprintf(o, "%s:%s:%s:%d:%d:%d:%d:%s", id->instrument, id->run, id->flowcell, id->lane, id->tile, id->x, id->y, id->tag);
	o += sprintf(o, ";%f", exp(quality));
	*o++ = '\n';
	for (size_t i = 0; i < seq_len; i++)
		*o++ = po_nt_to_ascii((char) seq_nt[i]);
	if (fastq) {
		*o++ = '\n';
		*o++ = '+';
		*o++ = '\n';
		for (size_t i = 0; i < seq_len; i++)
			*o++ = (char) (33 + po_result_phred(seq_p[i]));
	}
	*o++ = '\n';
	return (size_t) (o - dst);
}

/* Formats every OK pair of a flat result set, in order.  ids[i] belongs to pair i. */
size_t po_format_flat(char *dst, int fastq, size_t n, const po_seq_identifier *ids, const uint8_t *status,
                      const double *quality, const int32_t *seq_len, const uint8_t *seq_nt, const double *seq_p,
                      int64_t seq_stride) {
	size_t total = 0;
	for (size_t i = 0; i < n; i++) {
		if (status[i] != PO_OK)
			continue;
		total += po_format_record(dst + total, fastq, &ids[i], quality[i], seq_nt + i * (size_t) seq_stride,
		                          seq_p ? seq_p + i * (size_t) seq_stride : NULL, (size_t) seq_len[i]);
	}
	return total;
}
