/* oracle_pss.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the pss-bam / fragkon / genome-kmer-count hot paths.
 * See oracle_pss.h for the rules on who may call this.  Parity status: PINNED
 * against the unmodified reference binaries (oracle/_ref) and the committed
 * golden fixtures (tests/golden/).
 *
 * The restatement keeps the reference's *string level* formulation on purpose
 * (window copy, upper-casing, reverse-complement of both strings, then the two
 * per-end loops) so that the CUDA path, which uses an algebraically different
 * formulation (packed 2-bit windows, strand handled by complementing the cell
 * index), is checked against something that is structurally close to the
 * reference rather than to itself.
 */
/* no _GNU_SOURCE: with it glibc >= 2.38 routes sscanf to __isoc23_sscanf, whose %i also accepts "0b..." --
 * the reference (sam-parse.c, no feature macros, -std=gnu17) does not get that variant. */
#include "oracle_pss.h"

#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define ORA_MAX_LINE   200000   /* sam-parse.h:8  MAX_LINE_LEN  */
#define ORA_FIELD_W    2047     /* sam-parse.h:10 MAX_FIELD_WIDTH */
#define ORA_MAX_ID     511      /* fasta-genome-io.h:8 */
#define ORA_MAX_SEQ    536870911UL /* fasta-genome-io.h:9 */

/* ------------------------------------------------------------------------ */
/* FASTA                                                                      */
/* ------------------------------------------------------------------------ */

typedef struct bytesrc {
    const unsigned char *p;
    size_t               n, i;
} bytesrc;

static int src_get(bytesrc *s) { return s->i < s->n ? (int)s->p[s->i++] : EOF; }
static void src_unget(bytesrc *s, int c) { if (c != EOF && s->i > 0) s->i--; }

/* One record, following read_fasta (fasta-genome-io.c:105-148): the reference
 * keeps the byte in a plain `char`, so 0xFF compares equal to EOF; we keep
 * that by doing the comparisons on a char too. */
static int next_record(bytesrc *s, ora_contig *out, char *scratch)
{
    char   c;
    size_t i = 0;
    char   idbuf[ORA_MAX_ID + 1];

    c = (char)src_get(s);
    if (c != '>') {
        if (c == (char)EOF) return -1;
        /* A stream that does not start with '>' is outside the reference's
         * defined behaviour (uninitialised id, :135-147); stop here. */
        return -1;
    }
    c = (char)src_get(s);
    while (!isspace((unsigned char)c) && c != (char)EOF && i < ORA_MAX_ID) {
        idbuf[i++] = c;
        c = (char)src_get(s);
    }
    idbuf[i] = '\0';
    while (c != '\n' && c != (char)EOF) c = (char)src_get(s);

    i = 0;
    while (c != '>' && c != (char)EOF && i < ORA_MAX_SEQ) {
        if (!isspace((unsigned char)c)) scratch[i++] = (char)toupper((unsigned char)c);
        c = (char)src_get(s);
    }
    scratch[i] = '\0';
    if (c != (char)EOF) src_unget(s, c);

    out->id  = strdup(idbuf);
    out->seq = (char *)malloc(i + 1);
    memcpy(out->seq, scratch, i + 1);
    out->len = i;
    return 0;
}

static int contig_cmp(const void *a, const void *b)
{
    return strcmp(((const ora_contig *)a)->id, ((const ora_contig *)b)->id);
}

ora_genome *ora_genome_parse(const char *text, size_t len)
{
    bytesrc     s = { (const unsigned char *)text, len, 0 };
    ora_genome *g = (ora_genome *)calloc(1, sizeof *g);
    size_t      cap = 16;
    char       *scratch = (char *)malloc(len + 2);
    ora_contig  c;

    g->ctg = (ora_contig *)malloc(cap * sizeof *g->ctg);
    while (next_record(&s, &c, scratch) == 0) {
        if (g->n == cap) {
            cap *= 2;
            g->ctg = (ora_contig *)realloc(g->ctg, cap * sizeof *g->ctg);
        }
        g->ctg[g->n++] = c;
    }
    free(scratch);
    qsort(g->ctg, g->n, sizeof *g->ctg, contig_cmp);   /* fasta-genome-io.c:236 */
    return g;
}

/* Same genome from contigs already split out of their FASTA text (used for
 * multi-gigabase test genomes, where formatting and re-parsing FASTA text
 * would only cost time): bytes are upper-cased like read_fasta does
 * (fasta-genome-io.c:127) and the contigs sorted by id (:236). */
ora_genome *ora_genome_from_contigs(const char *const *ids, const char *const *seqs, const size_t *lens, size_t n)
{
    ora_genome *g = (ora_genome *)calloc(1, sizeof *g);
    size_t      i, k;
    g->ctg = (ora_contig *)malloc((n ? n : 1) * sizeof *g->ctg);
    g->n = n;
    for (i = 0; i < n; i++) {
        g->ctg[i].id = strdup(ids[i]);
        g->ctg[i].seq = (char *)malloc(lens[i] + 1);
        for (k = 0; k < lens[i]; k++) g->ctg[i].seq[k] = (char)toupper((unsigned char)seqs[i][k]);
        g->ctg[i].seq[lens[i]] = '\0';
        g->ctg[i].len = lens[i];
    }
    qsort(g->ctg, g->n, sizeof *g->ctg, contig_cmp);
    return g;
}

/* fasta-genome-io.c:6-15 is_gz: name ends in ".gz" */
static int name_is_gz(const char *fn)
{
    size_t n = strlen(fn);
    return n >= 3 && fn[n - 3] == '.' && fn[n - 2] == 'g' && fn[n - 1] == 'z';
}

ora_genome *ora_genome_load(const char *fasta_path)
{
    size_t cap = 1 << 20, n = 0;
    char  *buf = (char *)malloc(cap);
    ora_genome *g;

    if (name_is_gz(fasta_path)) {
        gzFile z = gzopen(fasta_path, "r");
        int    r;
        if (!z) { free(buf); return NULL; }
        while ((r = gzread(z, buf + n, (unsigned)(cap - n))) > 0) {
            n += (size_t)r;
            if (n == cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
        }
        gzclose(z);
    } else {
        FILE  *f = fopen(fasta_path, "r");
        size_t r;
        if (!f) { free(buf); return NULL; }
        while ((r = fread(buf + n, 1, cap - n, f)) > 0) {
            n += r;
            if (n == cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
        }
        fclose(f);
    }
    g = ora_genome_parse(buf, n);
    free(buf);
    return g;
}

void ora_genome_free(ora_genome *g)
{
    size_t i;
    if (!g) return;
    for (i = 0; i < g->n; i++) { free(g->ctg[i].id); free(g->ctg[i].seq); }
    free(g->ctg);
    free(g);
}

long ora_find_contig(const ora_genome *g, const char *id)
{
    size_t lo = 0, hi = g->n;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        int    c = strcmp(id, g->ctg[mid].id);
        if (c == 0) return (long)mid;
        if (c < 0) hi = mid; else lo = mid + 1;
    }
    return -1;
}

/* ------------------------------------------------------------------------ */
/* SAM record (sam-parse.h:20-56, only the fields the hot path reads)         */
/* ------------------------------------------------------------------------ */

typedef struct ora_rec {
    char          qname[ORA_FIELD_W + 1];
    unsigned int  flag;
    char          rname[ORA_FIELD_W + 1];
    unsigned long pos;
    unsigned int  mapq;
    char          cigar[ORA_FIELD_W + 1];
    char          mrnm[ORA_FIELD_W + 1];
    unsigned int  mpos;
    int           isize;
    int           seq_len;
    char          seq[ORA_FIELD_W + 1];
    char          qual[ORA_FIELD_W + 1];
} ora_rec;

#define FL_PAIRED 0x1u
#define FL_PROPER 0x2u
#define FL_UNMAP  0x4u
#define FL_MUNMAP 0x8u
#define FL_REV    0x10u
#define FL_READ1  0x40u
#define FL_READ2  0x80u
#define FL_SECOND 0x100u
#define FL_QCFAIL 0x200u
#define FL_DUP    0x400u
#define FL_SUPP   0x800u

/* sam-parse.c:10-91 line2saml: 11 whitespace-separated tokens via one sscanf,
 * strlen(seq)==strlen(qual), and isize := strlen(seq) for unpaired records
 * (:66-68).  Returns 0 ok / 1 bad. */
static int parse_line(const char *line, ora_rec *r)
{
    int got;
    {   /* The reference scans with unbounded "%s" into 2048-byte fields
         * (sam-parse.h:10,21-48): a token longer than 2047 bytes overflows
         * them.  Not reproducible -> the line is reported as undefined.  Only
         * the first eleven white-space delimited runs can reach a field. */
        const unsigned char *p = (const unsigned char *)line;
        int runs = 0;
        while (*p && runs < 11) {
            size_t n = 0;
            while (*p && isspace(*p)) p++;
            while (*p && !isspace(*p)) { p++; n++; }
            if (n > ORA_FIELD_W) return 2;
            if (n) runs++;
        }
    }
    got = sscanf(line,
                     "%2047s %u %2047s %lu %u %2047s %2047s %u %i %2047s %2047s",
                     r->qname, &r->flag, r->rname, &r->pos, &r->mapq, r->cigar,
                     r->mrnm, &r->mpos, &r->isize, r->seq, r->qual);
    if (got < 11) return 1;
    if (strlen(r->seq) != strlen(r->qual)) return 1;
    r->seq_len = (int)strlen(r->seq);
    if (!(r->flag & FL_PAIRED)) r->isize = r->seq_len;
    return 0;
}

/* Test hook: the conversions of one line, for fuzzing other sscanf
 * restatements against glibc's.  Returns 0 ok / 1 line2saml fails / 2 undefined.
 * isize is reported BEFORE the unpaired override of sam-parse.c:66-68. */
int ora_parse_line(const char *line, unsigned int *flag, unsigned long *pos, unsigned int *mapq,
                   int *isize, char *rname, char *cigar, char *seq)
{
    ora_rec *r = (ora_rec *)calloc(1, sizeof *r);
    int      rc;
    {
        const unsigned char *p = (const unsigned char *)line;
        int runs = 0;
        while (*p && runs < 11) {
            size_t n = 0;
            while (*p && isspace(*p)) p++;
            while (*p && !isspace(*p)) { p++; n++; }
            if (n > ORA_FIELD_W) { free(r); return 2; }
            if (n) runs++;
        }
    }
    rc = sscanf(line, "%2047s %u %2047s %lu %u %2047s %2047s %u %i %2047s %2047s",
                r->qname, &r->flag, r->rname, &r->pos, &r->mapq, r->cigar,
                r->mrnm, &r->mpos, &r->isize, r->seq, r->qual);
    if (rc < 11 || strlen(r->seq) != strlen(r->qual)) { free(r); return 1; }
    *flag = r->flag; *pos = r->pos; *mapq = r->mapq; *isize = r->isize;
    strcpy(rname, r->rname); strcpy(cigar, r->cigar); strcpy(seq, r->seq);
    free(r);
    return 0;
}

/* pss-bam.c:113-123 / fragkon.c:68-78 cigar_ok: CIGAR must be "<n>M" */
static int cigar_is_nM(int n, const char *cigar)
{
    char buf[32];
    snprintf(buf, sizeof buf, "%dM", n);
    return strcmp(buf, cigar) == 0;
}

/* pss-bam.c:60-79 do_revcomp / fragkon.c:27-46 do_rvcmp */
static void revcomp(const char *in, char *out, size_t n)
{
    size_t i;
    for (i = 0; i < n; i++) {
        char b = in[n - 1 - i], o;
        switch (b) {
        case 'A': case 'a': o = 'T'; break;
        case 'C': case 'c': o = 'G'; break;
        case 'G': case 'g': o = 'C'; break;
        case 'T': case 't': o = 'A'; break;
        default:            o = b;   break;
        }
        out[i] = o;
    }
}

static int base_code(char c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default:  return -1;
    }
}

/* pss-bam.c:169-189 add_ctx_counts */
static void add_ctx(uint64_t *tab, char first_ctx, char second_ctx)
{
    int c;
    if ((c = base_code(second_ctx)) >= 0) tab[0 * 16 + 5 * c] += 1;   /* row 0: 2 away   */
    if ((c = base_code(first_ctx))  >= 0) tab[1 * 16 + 5 * c] += 1;   /* row 1: adjacent */
}

/* pss-bam.c:197-257 add_fwd_counts: pair = {read, ref}, column 4*read+ref */
static void add_fwd(uint64_t *tab, const char *g, const char *r, int R)
{
    int i;
    for (i = 0; i < R; i++) {
        int a = base_code(r[i]), b = base_code(g[2 + i]);
        if (a >= 0 && b >= 0) tab[(i + 2) * 16 + 4 * a + b] += 1;
    }
}

/* pss-bam.c:266-326 add_rev_counts */
static void add_rev(uint64_t *tab, const char *g, const char *r, int n, int R)
{
    int i;
    for (i = 0; i < R; i++) {
        int a = base_code(r[n - 1 - i]), b = base_code(g[n + 1 - i]);
        if (a >= 0 && b >= 0) tab[(i + 2) * 16 + 4 * a + b] += 1;
    }
}

void ora_pss_default_params(ora_pss_params *p)
{
    p->region_len  = 15;
    p->min_len     = 0;
    p->max_len     = 250000000UL;
    p->min_mq      = 0;
    p->up_ctx      = "ACGT";
    p->down_ctx    = "ACGT";
    p->merged_only = 0;
}

/* pss-bam.c:390-496 process_aln */
static int pss_process(const ora_genome *g, const ora_pss_params *P, ora_rec *r,
                       uint64_t *fwd, uint64_t *rev, char *gbuf, char *rcg, char *rcr)
{
    const int R = P->region_len;
    long ci = ora_find_contig(g, r->rname);
    const ora_contig *ref;
    int  n;
    long s, e;
    size_t k;

    if (ci < 0) return ORA_NO_CONTIG;
    ref = &g->ctg[ci];
    /* `ref->len-1` wraps for an empty contig (:408) and the window copy then
     * reads past the string: undefined. */
    if (ref->len == 0) return ORA_UNDEFINED;
    /* :402 puts three char[n+5] arrays on the stack before any filter runs;
     * a paired record with a huge TLEN overflows the stack (SIGSEGV observed
     * at TLEN=5e7).  Beyond 1e6 the record is reported as
     * undefined; abs(INT_MIN) is negative and lands here too. */
    if (r->isize == INT_MIN || abs(r->isize) > 1000000) return ORA_UNDEFINED;

    n = abs(r->isize);                       /* :401 */
    s = (long)(r->pos - 1);                  /* :403 */
    e = s + n - 1;                           /* :404 */

    if (s - 2 < 0) return ORA_FILTERED;                                    /* :407 */
    if ((unsigned long)(e + 2) > (unsigned long)(ref->len - 1)) return ORA_FILTERED; /* :408 */
    if (r->mapq < (unsigned int)P->min_mq) return ORA_FILTERED;            /* :409 */
    if (!((unsigned long)n >= P->min_len && (unsigned long)n <= P->max_len && n >= R))
        return ORA_FILTERED;                                               /* :96-103 */
    if (!cigar_is_nM(n, r->cigar)) return ORA_FILTERED;                    /* :411 */
    if (r->flag & (FL_UNMAP | FL_SECOND | FL_QCFAIL | FL_DUP | FL_SUPP)) return ORA_FILTERED;
    if (P->merged_only && (r->flag & FL_PAIRED)) return ORA_FILTERED;      /* :417 */

    /* Paired reads take n from |TLEN|; if SEQ is shorter than that the
     * reference indexes past SEQ's terminator into bytes left over from
     * earlier records.  Not reproducible -> dropped and reported. */
    if ((r->flag & FL_PAIRED) && r->seq_len < n) return ORA_UNDEFINED;

    memcpy(gbuf, ref->seq + (s - 2), (size_t)n + 4);   /* :423 */
    gbuf[n + 4] = '\0';
    for (k = 0; gbuf[k]; k++) gbuf[k] = (char)toupper((unsigned char)gbuf[k]);     /* :424 */
    for (k = 0; r->seq[k]; k++) r->seq[k] = (char)toupper((unsigned char)r->seq[k]); /* :425 */

    if (!(r->flag & FL_PAIRED)) {                                          /* :428 */
        if (r->flag & FL_REV) {
            revcomp(gbuf, rcg, (size_t)n + 4);
            rcg[n + 4] = '\0';
            if (strchr(P->up_ctx, rcg[1]) && strchr(P->down_ctx, rcg[n + 2])) {
                revcomp(r->seq, rcr, (size_t)n);
                add_ctx(fwd, rcg[1], rcg[0]);
                add_ctx(rev, rcg[n + 2], rcg[n + 3]);
                add_fwd(fwd, rcg, rcr, R);
                add_rev(rev, rcg, rcr, n, R);
                return ORA_COUNTED;
            }
        } else if (strchr(P->up_ctx, gbuf[1]) && strchr(P->down_ctx, gbuf[n + 2])) {
            add_ctx(fwd, gbuf[1], gbuf[0]);
            add_ctx(rev, gbuf[n + 2], gbuf[n + 3]);
            add_fwd(fwd, gbuf, r->seq, R);
            add_rev(rev, gbuf, r->seq, n, R);
            return ORA_COUNTED;
        }
    } else if ((r->flag & FL_PROPER) && !(r->flag & FL_MUNMAP)) {          /* :450-452 */
        if (r->flag & FL_REV) {
            revcomp(gbuf, rcg, (size_t)n + 4);
            rcg[n + 4] = '\0';
            if ((r->flag & FL_READ1) && strchr(P->up_ctx, rcg[1])) {       /* :460 */
                revcomp(r->seq, rcr, (size_t)n);
                add_ctx(fwd, rcg[1], rcg[0]);
                add_fwd(fwd, rcg, rcr, R);
                return ORA_COUNTED;
            } else if ((r->flag & FL_READ2) && strchr(P->down_ctx, rcg[n + 2])) { /* :471 */
                revcomp(r->seq, rcr, (size_t)n);
                add_ctx(rev, rcg[n + 2], rcg[n + 3]);
                add_rev(rev, rcg, rcr, n, R);
                return ORA_COUNTED;
            }
        } else if ((r->flag & FL_READ1) && strchr(P->up_ctx, gbuf[1])) {   /* :482 */
            add_ctx(fwd, gbuf[1], gbuf[0]);
            add_fwd(fwd, gbuf, r->seq, R);
            return ORA_COUNTED;
        } else if ((r->flag & FL_READ2) && strchr(P->down_ctx, gbuf[n + 2])) { /* :488 */
            add_ctx(rev, gbuf[n + 2], gbuf[n + 3]);
            add_rev(rev, gbuf, r->seq, n, R);
            return ORA_COUNTED;
        }
    }
    return ORA_FILTERED;
}

/* Iterate the byte block the way `fgets(buf, MAX_LINE_LEN+1, fp)` does
 * (pss-bam.c:764): a "line" ends after '\n' or after 200000 bytes. */
typedef struct line_iter {
    const char *p;
    size_t      n, i;
    char       *buf;
} line_iter;

static int next_line(line_iter *it)
{
    size_t k = 0;
    if (it->i >= it->n) return 0;
    while (it->i < it->n && k < ORA_MAX_LINE) {
        char c = it->p[it->i++];
        it->buf[k++] = c;
        if (c == '\n') break;
    }
    it->buf[k] = '\0';
    return 1;
}

static void tally_status(ora_stats *st, int code)
{
    if (!st) return;
    st->lines++;
    switch (code) {
    case ORA_COUNTED:    st->counted++;    break;
    case ORA_NO_CONTIG:  st->no_contig++;  break;
    case ORA_FILTERED:   st->filtered++;   break;
    case ORA_PARSE_FAIL: st->parse_fail++; break;
    default:             st->undefined++;  break;
    }
}

uint64_t ora_pss_tally(const ora_genome *g, const char *sam, size_t sam_len,
                       const ora_pss_params *P,
                       uint64_t *fwd, uint64_t *rev,
                       int8_t *status, size_t status_cap, ora_stats *st)
{
    line_iter it = { sam, sam_len, 0, (char *)malloc(ORA_MAX_LINE + 2) };
    ora_rec  *r = (ora_rec *)calloc(1, sizeof *r);
    /* paired reads take n from |TLEN| (int) so size scratch generously;
     * anything larger than this is dropped before the buffers are touched
     * because no contig can hold it (ORA_MAX_SEQ). */
    size_t    cap = (size_t)ORA_MAX_SEQ + 8;
    char     *gbuf = NULL, *rcg = NULL, *rcr = NULL;
    size_t    have = 0;
    uint64_t  nlines = 0;

    if (st) memset(st, 0, sizeof *st);
    while (next_line(&it)) {
        int code;
        int pr = parse_line(it.buf, r);
        if (pr) {
            code = pr == 2 ? ORA_UNDEFINED : ORA_PARSE_FAIL;
        } else {
            size_t need = (size_t)abs(r->isize) + 8;
            if (need > cap) need = cap;
            if (need > have) {
                have = need < 4096 ? 4096 : need;
                gbuf = (char *)realloc(gbuf, have);
                rcg  = (char *)realloc(rcg, have);
                rcr  = (char *)realloc(rcr, have);
            }
            code = pss_process(g, P, r, fwd, rev, gbuf, rcg, rcr);
        }
        if (status && nlines < status_cap) status[nlines] = (int8_t)code;
        tally_status(st, code);
        nlines++;
    }
    free(gbuf); free(rcg); free(rcr); free(r); free(it.buf);
    return nlines;
}

void ora_pss_rates(const uint64_t *counts, int R, double *rates)
{
    /* 12 off-diagonal pairs in print order AC AG AT CA CG CT GA GC GT TA TC TG
     * (pss-bam.c:515-526): numerator column, denominator = column sum of ref */
    static const int num[12] = { 1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14 };
    int i, j;
    for (i = 0; i < R * 12; i++) rates[i] = 0.0;
    for (i = 0; i < R; i++) {
        const uint64_t *c = counts + (size_t)(i + 2) * 16;
        double nref[4];
        for (j = 0; j < 4; j++)
            nref[j] = (double)(c[j] + c[4 + j] + c[8 + j] + c[12 + j]);   /* :508-511 */
        if (nref[0] == 0 || nref[1] == 0 || nref[2] == 0 || nref[3] == 0) continue; /* :512 */
        for (j = 0; j < 12; j++)
            rates[i * 12 + j] = (double)c[num[j]] / nref[num[j] & 3];
    }
}

int ora_pss_write_counts(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                         const uint64_t *fwd, const uint64_t *rev, int R)
{
    char  fn[ORA_FIELD_W];
    FILE *fp;
    int   i, j;
    snprintf(fn, sizeof fn, "%s.pss.counts.txt", out_prefix);
    if (!(fp = fopen(fn, "w"))) return 1;
    fprintf(fp, "### pss-bam.c v1.2.1:\n### FASTA: %s\n### BAM: %s\n### OUT: %s\n",
            fasta_fn, bam_fn, fn);
    fputs("### Format of table:\n"
          "### Counts of how often a read base and genome base were seen at\n"
          "### each position in the aligned reads.\n"
          "### First base is what was seen in the read.\n"
          "### Second base is what was in the genome at that position.\n"
          "### POS AA AC AG AT CA CC CG CT GA GC GG GT TA TC TG TT\n"
          "### Forward read substitution counts and base context\n", fp);
    for (i = -2; i < R; i++) {
        fprintf(fp, "%d\t", i);
        for (j = 0; j < 16; j++) fprintf(fp, "%lu\t", (unsigned long)fwd[(i + 2) * 16 + j]);
        fputc('\n', fp);
    }
    fputs("\n\n### Reverse read substitution counts and base context\n", fp);
    for (i = R - 1; i >= 0; i--) {
        fprintf(fp, "%d\t", i);
        for (j = 0; j < 16; j++) fprintf(fp, "%lu\t", (unsigned long)rev[(i + 2) * 16 + j]);
        fputc('\n', fp);
    }
    for (i = 1; i < 3; i++) {              /* context rows, labelled 1 then 2 (:577-583) */
        fprintf(fp, "%d\t", i);
        for (j = 0; j < 16; j++) fprintf(fp, "%lu\t", (unsigned long)rev[(2 - i) * 16 + j]);
        fputc('\n', fp);
    }
    fclose(fp);
    return 0;
}

int ora_pss_write_rates(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                        const double *fr, const double *rr, int R)
{
    char  fn[ORA_FIELD_W];
    FILE *fp;
    int   i, j;
    snprintf(fn, sizeof fn, "%s.pss.rates.txt", out_prefix);
    if (!(fp = fopen(fn, "w"))) return 1;
    fprintf(fp, "### pss-bam.c v%s\n### FASTA: %s\n### BAM: %s\n### OUT: %s\n",
            "1.2.1", fasta_fn, bam_fn, fn);
    fputs("### Format of table:\n"
          "### Substitution rates for all possible nucleotide substitutions at\n"
          "### each position in the aligned reads.\n"
          "### First base is what was seen in the read.\n"
          "### Second base is what was in the genome at that position.\n"
          "### POS AC AG AT CA CG CT GA GC GT TA TC TG\n"
          "### Forward read substitution rates\n", fp);
    for (i = 0; i < R; i++) {
        fprintf(fp, "%d\t", i);
        for (j = 0; j < 12; j++) fprintf(fp, "%.5e\t", fr[i * 12 + j]);
        fputc('\n', fp);
    }
    fputs("\n\n### Reverse read substitution rates\n", fp);
    for (i = R - 1; i >= 0; i--) {
        fprintf(fp, "%d\t", i);
        for (j = 0; j < 12; j++) fprintf(fp, "%.5e\t", rr[i * 12 + j]);
        fputc('\n', fp);
    }
    fclose(fp);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* k-mers (kmer.c)                                                            */
/* ------------------------------------------------------------------------ */

int ora_kmer2inx(const char *kmer, size_t klen, uint64_t *inx)
{
    uint64_t v = 0;
    size_t   i;
    for (i = 0; i < klen; i++) {
        int c = base_code((char)toupper((unsigned char)kmer[i]));   /* kmer.c:193 */
        if (c < 0) return 0;
        v = (v << 2) | (uint64_t)c;
    }
    *inx = v;
    return 1;
}

/* add_to_ksp (kmer.c:43-110) as a flat histogram: the array+trie structure
 * indexes exactly the set of all-ACGT k-mers, MSB-first; any other character
 * (including the string terminator) rejects the k-mer. */
static int kmer_add(const char *p, int k, uint64_t *tab)
{
    uint64_t inx;
    if (!ora_kmer2inx(p, (size_t)k, &inx)) return -1;
    tab[inx] += 1;
    return 0;
}

void ora_kmer_spectrum(const ora_genome *g, int k, uint64_t *counts)
{
    size_t c, i;
    for (c = 0; c < g->n; c++) {
        const ora_contig *s = &g->ctg[c];
        /* genome-kmer-count.c:71 computes len-k+1 in size_t; for len < k-1
         * that underflows (undefined) -- such contigs contribute nothing here. */
        for (i = 0; i + (size_t)k <= s->len; i++) kmer_add(s->seq + i, k, counts);
    }
}

static void inx2kmer(uint64_t inx, int k, char *out)
{
    int i;
    out[k] = '\0';
    for (i = 0; i < k; i++) {               /* genome-kmer-count.c:85-115 */
        out[k - 1 - i] = "ACGT"[inx & 3];
        inx >>= 2;
    }
}

static unsigned int clamp_u32(uint64_t v) { return v > UINT_MAX ? UINT_MAX : (unsigned int)v; }

int ora_kmer_spectrum_write(void *FILE_out, size_t n_seqs, int k, const uint64_t *counts)
{
    FILE    *fp = (FILE *)FILE_out;
    uint64_t i, n = 1ULL << (2 * k);
    char     km[40];
    fprintf(fp, "Parsed input genome. Found %lu sequences.\n", (unsigned long)n_seqs);
    for (i = 0; i < n; i++) {
        inx2kmer(i, k, km);
        fprintf(fp, "%s\t%u\n", km, clamp_u32(counts[i]));
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* fragkon (fragkon.c)                                                        */
/* ------------------------------------------------------------------------ */

void ora_fk_default_params(ora_fk_params *p)
{
    p->klen = 8; p->min_len = 0; p->max_len = 250000000UL; p->min_mq = 0; p->merged_only = 0;
}

/* Character at genome index i as the reference would see it.  Indices in
 * [0,len) are the contig; index len is the string terminator (rejects the
 * k-mer).  Anything else is outside the allocation (fragkon.c:137
 * `aln_start-ok >= 0` is an unsigned tautology, so reads with fewer than K/2
 * bases to the left of the alignment are NOT filtered): the reference then
 * reads the allocator's chunk header, whose last byte (index -1) is the top
 * byte of a size field and therefore 0 under glibc.  We return '\0' for every
 * out-of-range index, which reproduces the observed behaviour: forward reads
 * lose the 5' k-mer and keep the 3' one; reverse reads lose both, because the
 * strncpy at fragkon.c:101-103 stops at that 0 and zero-fills the rest. */
static char gchar(const ora_contig *ref, long i)
{
    if (i < 0 || (unsigned long)i >= ref->len) return '\0';
    return ref->seq[i];
}

/* fragkon.c:122-216 process_aln */
static int fk_process(const ora_genome *g, const ora_fk_params *P, const ora_rec *r,
                      uint64_t *fp, uint64_t *tp)
{
    const int K = P->klen;
    long ci = ora_find_contig(g, r->rname);
    const ora_contig *ref;
    unsigned long s, e;
    unsigned int  ok = (unsigned int)(K / 2), ik = (unsigned int)K - ok;   /* :134-135 */
    int   n = r->seq_len;                                                  /* :130 */
    char  w5[64], w3[64];
    int   j, a5, a3;

    if (ci < 0) return ORA_NO_CONTIG;
    ref = &g->ctg[ci];
    if (ref->len == 0) return ORA_UNDEFINED;   /* `ref->len-1` wraps (:138) */
    s = r->pos - 1;                                                        /* :129 */
    e = s + (unsigned long)n - 1;                                          /* :130 */

    if (!(e + (unsigned long)(K / 2) <= ref->len - 1)) return ORA_FILTERED;       /* :138 */
    if (!(r->mapq >= (unsigned int)P->min_mq)) return ORA_FILTERED;               /* :139 */
    if (!((unsigned long)n >= P->min_len && (unsigned long)n <= P->max_len)) return ORA_FILTERED;
    if (!cigar_is_nM(n, r->cigar)) return ORA_FILTERED;                           /* :141 */
    if (r->flag & (FL_UNMAP | FL_SECOND | FL_QCFAIL | FL_DUP | FL_SUPP)) return ORA_FILTERED;

    if (r->flag & FL_REV) {
        /* sub = strncpy of G[s-ok, s-ok+n+K) (:156, stops at a terminator and
         * zero-fills); rc = RC(sub) (:160); 5' = rc[0..K) (:164),
         * 3' = rc[ok+n-ik .. +K) (:167). */
        size_t L = (size_t)n + (size_t)K, q;
        char  *sub = (char *)malloc(L + 1), *rc = (char *)malloc(L + 1);
        int    hit_nul = 0;
        for (q = 0; q < L; q++) {
            char c = hit_nul ? '\0' : gchar(ref, (long)s - (long)ok + (long)q);
            if (c == '\0') hit_nul = 1;
            sub[q] = c;
        }
        revcomp(sub, rc, L);
        for (j = 0; j < K; j++) w5[j] = rc[j];
        for (j = 0; j < K; j++) w3[j] = rc[(size_t)ok + (size_t)n - (size_t)ik + (size_t)j];
        free(sub); free(rc);
    } else {
        for (j = 0; j < K; j++) w5[j] = gchar(ref, (long)s - (long)ok + j);          /* :176 */
        for (j = 0; j < K; j++) w3[j] = gchar(ref, (long)s + n - (long)ik + j);      /* :177 */
    }
    w5[K] = w3[K] = '\0';

    if (!(r->flag & FL_PAIRED)) {                                                   /* :149 */
        a5 = kmer_add(w5, K, fp);
        a3 = kmer_add(w3, K, tp);
        return (a5 == 0 && a3 == 0) ? ORA_COUNTED : ORA_FILTERED;
    }
    if (!P->merged_only && (r->flag & FL_PROPER) && !(r->flag & FL_MUNMAP)) {       /* :187-190 */
        if (r->flag & FL_READ1) return kmer_add(w5, K, fp) == 0 ? ORA_COUNTED : ORA_FILTERED;
        if (r->flag & FL_READ2) return kmer_add(w3, K, tp) == 0 ? ORA_COUNTED : ORA_FILTERED;
    }
    return ORA_FILTERED;
}

uint64_t ora_fragkon_tally(const ora_genome *g, const char *sam, size_t sam_len,
                           const ora_fk_params *P, uint64_t *fp, uint64_t *tp,
                           int8_t *status, size_t status_cap, ora_stats *st)
{
    line_iter it = { sam, sam_len, 0, (char *)malloc(ORA_MAX_LINE + 2) };
    ora_rec  *r = (ora_rec *)calloc(1, sizeof *r);
    uint64_t  nlines = 0;

    if (st) memset(st, 0, sizeof *st);
    while (next_line(&it)) {
        int pr = parse_line(it.buf, r);
        int code = pr == 2 ? ORA_UNDEFINED : pr ? ORA_PARSE_FAIL : fk_process(g, P, r, fp, tp);
        if (status && nlines < status_cap) status[nlines] = (int8_t)code;
        tally_status(st, code);
        nlines++;
    }
    free(r); free(it.buf);
    return nlines;
}

int ora_fragkon_write(void *FILE_out, const char *fasta_fn, const char *bam_fn,
                      int K, const uint64_t *fp, const uint64_t *tp)
{
    FILE    *f = (FILE *)FILE_out;
    uint64_t i, n = 1ULL << (2 * K);
    char     km[40];
    fprintf(f, "### fragkon.c v0.3\n### %s\n### %s\n", fasta_fn, bam_fn);
    fprintf(f, "# KMER\t5' CONTEXT COUNTS\t3' CONTEXT COUNTS\n");
    for (i = 0; i < n; i++) {
        inx2kmer(i, K, km);
        fprintf(f, "%s\t%u\t%u\n", km, clamp_u32(fp[i]), clamp_u32(tp[i]));
    }
    return 0;
}
