/* pss_fasta.c -- block-buffered FASTA reader behind the fasta-genome-io.h API. */
#include "pss_fasta.h"

#define PSS_FA_BLOCK (4u << 20)

/* ---- byte source: one refill per 4 MiB instead of one libc call per byte ---- */
typedef struct blockreader {
    FILE          *fp;
    gzFile         gz;
    unsigned char *buf;
    size_t         have, at;
    int            done;
} blockreader;

static blockreader *br_new(FILE *fp, gzFile gz)
{
    blockreader *r = (blockreader *)calloc(1, sizeof *r);
    r->fp = fp;
    r->gz = gz;
    r->buf = (unsigned char *)malloc(PSS_FA_BLOCK);
    return r;
}
static void br_free(blockreader *r)
{
    if (r) { free(r->buf); free(r); }
}
/* next byte without consuming it, or -1 at the end.  A 0xFF byte ends the
 * input too: the reference keeps fgetc()'s result in a plain char and compares
 * it with EOF (fasta-genome-io.c:106,120). */
static int br_peek(blockreader *r)
{
    if (r->at == r->have) {
        int got = 0;
        if (r->done) return -1;
        if (r->gz) got = gzread(r->gz, r->buf, PSS_FA_BLOCK);
        else got = (int)fread(r->buf, 1, PSS_FA_BLOCK, r->fp);
        r->at = 0;
        r->have = got > 0 ? (size_t)got : 0;
        if (r->have == 0) { r->done = 1; return -1; }
    }
    if (r->buf[r->at] == 0xFF) { r->done = 1; r->have = r->at; return -1; }
    return r->buf[r->at];
}
static void br_skip(blockreader *r) { r->at++; }

/* growable output string */
typedef struct strbuf { char *p; size_t n, cap; } strbuf;
static void sb_reserve(strbuf *s, size_t extra)
{
    if (s->n + extra + 1 > s->cap) {
        size_t c = s->cap ? s->cap : (1u << 20);
        while (c < s->n + extra + 1) c *= 2;
        s->p = (char *)realloc(s->p, c);
        s->cap = c;
    }
}

/* One record from the reader into *seq.  0 ok, -1 no more records. */
static int next_record(blockreader *r, Seq *seq)
{
    int    c;
    size_t i = 0;
    strbuf s = { NULL, 0, 0 };

    /* defined input starts each record with '>'; anything before the first one is skipped */
    while ((c = br_peek(r)) >= 0 && c != '>') br_skip(r);
    if (c < 0) return -1;
    br_skip(r);
    while ((c = br_peek(r)) >= 0 && !isspace(c)) {       /* id */
        if (i < MAX_ID_LEN) seq->id[i++] = (char)c;
        br_skip(r);
    }
    seq->id[i] = '\0';
    while ((c = br_peek(r)) >= 0 && c != '\n') br_skip(r);   /* rest of the header line */

    /* sequence: a line at a time.  The common line (bases only, then '\n') is found with memchr and copied
     * with a branch-free upper-casing loop the compiler vectorises; lines holding anything else (blanks, '\r',
     * a '>' in the middle, the 0xFF end marker) take the byte loop. */
    for (;;) {
        size_t avail, seg, k;
        const unsigned char *p, *nl;
        c = br_peek(r);
        if (c < 0 || c == '>') break;
        if (s.n >= MAX_SEQ_LEN) {                          /* truncate like the reference, then resynchronise */
            fprintf(stderr, "%s is truncated to %d\n", seq->id, MAX_SEQ_LEN);
            while ((c = br_peek(r)) >= 0 && c != '>') br_skip(r);
            break;
        }
        p = r->buf + r->at;
        avail = r->have - r->at;
        if (avail > MAX_SEQ_LEN - s.n) avail = MAX_SEQ_LEN - s.n;
        nl = (const unsigned char *)memchr(p, '\n', avail);
        seg = nl ? (size_t)(nl - p) : avail;
        sb_reserve(&s, seg);
        {
            unsigned odd = 0;                              /* any byte that is not plain sequence? */
            char *out = s.p + s.n;
            for (k = 0; k < seg; k++) {
                const unsigned char ch = p[k];
                odd |= (unsigned)(ch <= ' ') | (unsigned)(ch == '>') | (unsigned)(ch == 0xFF);
                out[k] = (char)(ch - (((unsigned char)(ch - 'a') < 26u) << 5));
            }
            if (!odd) {
                s.n += seg;
                r->at += seg + (nl ? 1 : 0);               /* the newline itself is white space: dropped */
                continue;
            }
        }
        for (k = 0; k < seg; k++) {                        /* rare line: byte by byte, stop at '>' / 0xFF */
            const unsigned char ch = p[k];
            if (ch == '>' || ch == 0xFF) break;
            if (!isspace(ch)) s.p[s.n++] = (char)toupper(ch);
        }
        r->at += k;
        if (k == seg && nl) r->at++;
    }
    sb_reserve(&s, 0);
    s.p[s.n] = '\0';
    seq->seq = (char *)realloc(s.p, s.n + 1);
    seq->len = s.n;
    return 0;
}

int is_gz(const char *fn)
{
    size_t n = strlen(fn);
    return n >= 3 && strcmp(fn + n - 3, ".gz") == 0;
}

FILE *fileOpen(const char *name, char access_mode[])
{
    FILE *f = fopen(name, access_mode);
    if (!f) fprintf(stderr, "Cannot open %s!\n", name);
    return f;
}

Fa_Src *init_fasta_src(const char fn[])
{
    Fa_Src *src = (Fa_Src *)calloc(1, sizeof *src);
    strncpy(src->fn, fn, MAX_FN_LEN);
    src->is_gz = is_gz(fn);
    if (src->is_gz) src->fagz = gzopen(fn, "r");
    else src->fafp = fileOpen(fn, (char *)"r");
    if (!src->fagz && !src->fafp) { free(src); return NULL; }
    src->seq_buffer = (char *)br_new(src->fafp, src->fagz);
    return src;
}

Seq *get_next_fa(Fa_Src *fa_source, Genome *genome)
{
    Seq *s;
    if (!fa_source) return NULL;
    /* like the reference (fasta-genome-io.c:60-83), the record is appended to genome->seqs, which the caller sized
     * for MAX_GENOME_SEQS pointers (init_genome does); a full table ends the iteration */
    if (genome && genome->n_seqs >= MAX_GENOME_SEQS) return NULL;
    s = (Seq *)calloc(1, sizeof *s);
    if (next_record((blockreader *)fa_source->seq_buffer, s) != 0) { free(s); return NULL; }
    fa_source->n++;
    if (genome) genome->seqs[genome->n_seqs++] = s;
    return s;
}

/* The two single-record entry points of the API read through a private
 * reader so that a caller may interleave them with its own stdio calls only at
 * record boundaries (which is how the reference's own loader used them). */
int read_fasta(FILE *fp, Seq *seq, char *seq_buffer)
{
    /* simple character loop that keeps the FILE position exact; not on any hot path */
    strbuf s = { NULL, 0, 0 };
    size_t i = 0;
    int    c;
    (void)seq_buffer;
    c = fgetc(fp);
    while (c != EOF && c != '>' && c != 0xFF) c = fgetc(fp);
    if (c != '>') return -1;
    while ((c = fgetc(fp)) != EOF && !isspace(c)) if (i < MAX_ID_LEN) seq->id[i++] = (char)c;
    seq->id[i] = '\0';
    while (c != EOF && c != '\n') c = fgetc(fp);
    while ((c = fgetc(fp)) != EOF && c != '>' && c != 0xFF) {
        if (isspace(c) || s.n >= MAX_SEQ_LEN) continue;
        sb_reserve(&s, 1);
        s.p[s.n++] = (char)toupper(c);
    }
    if (c == '>') ungetc(c, fp);
    sb_reserve(&s, 0);
    s.p[s.n] = '\0';
    seq->seq = s.p;
    seq->len = s.n;
    return 0;
}

int gzread_fasta(gzFile gzfp, Seq *seq, char *seq_buffer)
{
    strbuf s = { NULL, 0, 0 };
    size_t i = 0;
    int    c;
    (void)seq_buffer;
    c = gzgetc(gzfp);
    while (c != -1 && c != '>' && c != 0xFF) c = gzgetc(gzfp);
    if (c != '>') return -1;
    while ((c = gzgetc(gzfp)) != -1 && !isspace(c)) if (i < MAX_ID_LEN) seq->id[i++] = (char)c;
    seq->id[i] = '\0';
    while (c != -1 && c != '\n') c = gzgetc(gzfp);
    while ((c = gzgetc(gzfp)) != -1 && c != '>' && c != 0xFF) {
        if (isspace(c) || s.n >= MAX_SEQ_LEN) continue;
        sb_reserve(&s, 1);
        s.p[s.n++] = (char)toupper(c);
    }
    if (c == '>') gzungetc(c, gzfp);
    sb_reserve(&s, 0);
    s.p[s.n] = '\0';
    seq->seq = s.p;
    seq->len = s.n;
    return 0;
}

int close_fasta_src(Fa_Src *src)
{
    if (!src) return 0;
    br_free((blockreader *)src->seq_buffer);
    if (src->fagz) gzclose(src->fagz);
    if (src->fafp) fclose(src->fafp);
    free(src);
    return 0;
}

int chr_cmp(const void *v1, const void *v2)
{
    const Seq *a = *(const Seq *const *)v1, *b = *(const Seq *const *)v2;
    return strcmp(a->id, b->id);
}

Genome *init_genome(const char fn[])
{
    Fa_Src *src = init_fasta_src(fn);
    Genome *g;
    if (!src) return NULL;
    g = (Genome *)calloc(1, sizeof *g);
    /* the reference's table size (fasta-genome-io.c:226): 8 MB of address space, touched only as far as it is used */
    g->seqs = (Seq **)malloc((size_t)MAX_GENOME_SEQS * sizeof *g->seqs);
    g->dummy = (Seq *)calloc(1, sizeof *g->dummy);
    while (get_next_fa(src, g) != NULL) { }
    close_fasta_src(src);
    qsort(g->seqs, g->n_seqs, sizeof *g->seqs, chr_cmp);
    return g;
}

Seq *find_seq(Genome *genome, const char id[])
{
    Seq **hit, *key = genome->dummy;
    strncpy(key->id, id, MAX_ID_LEN);
    key->id[MAX_ID_LEN] = '\0';
    hit = (Seq **)bsearch(&key, genome->seqs, genome->n_seqs, sizeof *genome->seqs, chr_cmp);
    return hit ? *hit : NULL;
}

int destroy_seq(Seq *seq)
{
    if (seq) { free(seq->seq); free(seq); }
    return 0;
}

int destroy_genome(Genome *genome)
{
    size_t i;
    if (!genome) return 0;
    for (i = 0; i < genome->n_seqs; i++) destroy_seq(genome->seqs[i]);
    free(genome->seqs);
    free(genome->dummy);
    free(genome);
    return 0;
}
