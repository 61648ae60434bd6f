/* pss_synth.c -- seeded synthetic FASTA / SAM generators (test + bench
 * infrastructure, not product code).
 *
 * Everything is a pure function of (seed, index): contig c of the genome and
 * read i of the read set can be generated on any rank, in any order, in any
 * number of threads, and come out byte-identical.  The recipes follow
 * SURVEY.md section 8(d):
 *   config 1: uniform ACGT genome, fixed-length unpaired reads, CIGAR "<n>M",
 *             5' C->T / 3' G->A damage in molecule orientation, 0.2 % errors.
 *   config 2: human-like contigs with N runs and soft-masked (lower-case)
 *             stretches; variable-length reads; a mix of records the
 *             reference must reject (indel / soft-clip CIGARs, filtered
 *             flags, QUAL "*", unknown contigs, reads hanging over contig
 *             ends) and paired records (TLEN == n and TLEN != n).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct synth_reads_cfg {
    uint64_t seed;
    uint32_t min_len, max_len;      /* read length, uniform inclusive            */
    double   p_reverse;             /* FLAG 0x10                                  */
    double   p_indel;               /* CIGAR with an I or D op -> must be dropped */
    double   p_softclip;            /* CIGAR with S op         -> must be dropped */
    double   p_badflag;             /* one of 4/256/512/1024/2048 set             */
    double   p_paired;              /* flags 99/147/83/163, TLEN = +-n or != n    */
    double   p_qualstar;            /* QUAL "*"                -> parse reject    */
    double   p_unknown_contig;      /* RNAME not in the FASTA                     */
    double   p_edge;                /* alignment placed at/over a contig edge     */
    double   p_read_n;              /* one 'N' somewhere in the read              */
    double   damage5;               /* C->T at molecule 5' end: p = damage5*exp(-d/3) */
    double   damage3;               /* G->A at molecule 3' end                    */
    double   err;                   /* uniform substitution error per base        */
    uint32_t max_mapq;              /* MAPQ uniform 0..max_mapq                   */
    uint32_t with_tags;             /* append NM:i / MD:Z                         */
    uint64_t sorted_total;          /* > 0: coordinate-sorted output -- read idx of sorted_total sits at that fraction of the genome */
} synth_reads_cfg;

/* Thread count for the generators (torchrun exports OMP_NUM_THREADS=1 to every rank). */
void synth_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- counter based RNG -------------------------------------------------- */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline rng_t rng_at(uint64_t seed, uint64_t stream, uint64_t idx)
{
    rng_t r; r.s = mix64(seed ^ mix64(stream * 0x632be59bd9b4e019ULL + idx)); return r;
}
static inline uint64_t rng_u64(rng_t *r) { r->s += 0x9e3779b97f4a7c15ULL; return mix64(r->s); }
static inline double   rng_unit(rng_t *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return n ? (uint64_t)(rng_unit(r) * (double)n) : 0; }

/* ---- genome -------------------------------------------------------------- */
/* Fill out[0..len) with contig `cidx`.  Bases are uniform ACGT generated in
 * 64-base blocks (one RNG draw per 32 bases).  Then `n_frac` of the contig is
 * overwritten with runs of 'N' and `lower_frac` with lower-case runs. */
void synth_contig(uint64_t seed, uint32_t cidx, uint64_t len,
                  double n_frac, double lower_frac, char *out)
{
    static const char B[4] = { 'A', 'C', 'G', 'T' };
    uint64_t nblk = (len + 31) / 32;
    int64_t  b;
#pragma omp parallel for schedule(static)
    for (b = 0; b < (int64_t)nblk; b++) {
        rng_t    r = rng_at(seed, 1000 + cidx, (uint64_t)b);
        uint64_t v = rng_u64(&r);
        uint64_t i0 = (uint64_t)b * 32, i1 = i0 + 32 > len ? len : i0 + 32, i;
        for (i = i0; i < i1; i++) { out[i] = B[v & 3]; v >>= 2; }
    }
    {   /* N runs and soft-masked runs: sequential, cheap (few runs) */
        rng_t    r = rng_at(seed, 2000 + cidx, 0);
        uint64_t budget = (uint64_t)((double)len * n_frac), i;
        while (budget > 0 && len > 64) {
            uint64_t L = 50 + rng_below(&r, 5000);
            uint64_t at;
            if (L > budget) L = budget;
            if (L >= len) L = len - 1;
            at = rng_below(&r, len - L);
            memset(out + at, 'N', L);
            budget -= L;
        }
        budget = (uint64_t)((double)len * lower_frac);
        while (budget > 0 && len > 64) {
            uint64_t L = 100 + rng_below(&r, 3000);
            uint64_t at;
            if (L > budget) L = budget;
            if (L >= len) L = len - 1;
            at = rng_below(&r, len - L);
            for (i = at; i < at + L; i++) out[i] |= 0x20;
            budget -= L;
        }
    }
}

/* ---- reads ---------------------------------------------------------------- */
static inline char comp(char c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    default:  return c;
    }
}
static inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

static inline char *put_u64(char *p, uint64_t v)
{
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
static inline char *put_str(char *p, const char *s) { while (*s) *p++ = *s++; return p; }

#define SYNTH_MAX_READ 4096
#define SYNTH_MAX_REC  (4 * SYNTH_MAX_READ + 512)

/* One SAM record for read index `idx`; returns bytes written (incl. '\n'). */
static size_t one_read(const synth_reads_cfg *C, uint64_t idx,
                       const char *const *cseq, const uint64_t *clen,
                       const char *const *cname, const double *ccum, uint32_t nc,
                       char *out)
{
    rng_t    r = rng_at(C->seed, 7, idx);
    uint32_t n = C->min_len + (uint32_t)rng_below(&r, (uint64_t)(C->max_len - C->min_len) + 1);
    double   u = rng_unit(&r);
    const int sorted = C->sorted_total > 0;
    if (sorted) u = ((double)idx + 0.5) / (double)C->sorted_total;   /* contigs are picked by cumulative length: monotone in idx */
    uint32_t ci = 0;
    int      reverse, kind_indel, kind_clip, kind_badflag, kind_paired, kind_qstar, kind_unk, kind_edge;
    uint64_t L, start;     /* 0-based start */
    uint32_t flag = 0, mapq, i;
    int64_t  tlen = 0;
    char     mol[SYNTH_MAX_READ + 8], seq[SYNTH_MAX_READ + 8], refw[SYNTH_MAX_READ + 8];
    char     cigar[64], *p = out;
    uint32_t nm = 0;

    if (n > SYNTH_MAX_READ) n = SYNTH_MAX_READ;
    while (ci + 1 < nc && u >= ccum[ci]) ci++;
    L = clen[ci];

    reverse      = rng_unit(&r) < C->p_reverse;
    kind_indel   = rng_unit(&r) < C->p_indel;
    kind_clip    = rng_unit(&r) < C->p_softclip;
    kind_badflag = rng_unit(&r) < C->p_badflag;
    kind_paired  = rng_unit(&r) < C->p_paired;
    kind_qstar   = rng_unit(&r) < C->p_qualstar;
    kind_unk     = rng_unit(&r) < C->p_unknown_contig;
    kind_edge    = rng_unit(&r) < C->p_edge;
    mapq         = (uint32_t)rng_below(&r, (uint64_t)C->max_mapq + 1);

    if (L < (uint64_t)n + 64) {            /* contig too short for this read: shrink */
        n = L > 80 ? (uint32_t)(L - 64) : (L > 12 ? (uint32_t)(L - 8) : (uint32_t)(L > 1 ? L - 1 : 1));
    }
    if (kind_edge) {
        /* place so that the +-2 context falls on / over an edge */
        uint64_t k = rng_below(&r, 6);            /* 0..5 */
        if (rng_unit(&r) < 0.5) start = k;        /* pos = 1..6: pos 1,2 rejected, 3+ accepted */
        else                    start = L - n - k;  /* k = 0,1: e+2 > len-1 rejected; k >= 2 accepted */
        if (start + n > L) start = L - n;         /* tiny contigs */
    } else if (L >= (uint64_t)n + 64) {
        start = 32 + rng_below(&r, L - n - 64);
        if (sorted) {                      /* position within the contig follows idx as well (ties broken by the read length) */
            const double lo = ci ? ccum[ci - 1] : 0.0, hi = ccum[ci];
            const double f = hi > lo ? (u - lo) / (hi - lo) : 0.0;
            start = 32 + (uint64_t)(f * (double)(L - n - 64));
        }
    } else {
        start = rng_below(&r, L - n + 1);         /* tiny contig: anywhere, edges included */
    }

    /* reference window, upper-cased like the loader does */
    for (i = 0; i < n; i++) refw[i] = up(cseq[ci][start + i]);

    /* molecule orientation */
    if (reverse) for (i = 0; i < n; i++) mol[i] = comp(refw[n - 1 - i]);
    else         for (i = 0; i < n; i++) mol[i] = refw[i];
    for (i = 0; i < n; i++) {
        char c = mol[i];
        if (c == 'N') { c = "ACGT"[rng_below(&r, 4)]; }      /* reads over N runs carry real bases */
        {
            double p5 = C->damage5 * exp(-(double)i / 3.0);
            double p3 = C->damage3 * exp(-(double)(n - 1 - i) / 3.0);
            if (c == 'C' && rng_unit(&r) < p5) c = 'T';
            else if (c == 'G' && rng_unit(&r) < p3) c = 'A';
        }
        if (rng_unit(&r) < C->err) {
            const char *at = strchr("ACGT", c);
            c = "ACGT"[(rng_below(&r, 3) + 1 + (uint64_t)(at ? at - "ACGT" : 0)) & 3];
        }
        mol[i] = c;
    }
    if (rng_unit(&r) < C->p_read_n) mol[rng_below(&r, n)] = 'N';
    if (reverse) for (i = 0; i < n; i++) seq[i] = comp(mol[n - 1 - i]);
    else         for (i = 0; i < n; i++) seq[i] = mol[i];
    seq[n] = '\0';
    for (i = 0; i < n; i++) if (seq[i] != refw[i]) nm++;

    /* flags */
    if (reverse) flag |= 16;
    if (kind_paired) {
        int second = rng_unit(&r) < 0.5;
        int tlen_ok = rng_unit(&r) < 0.6;
        flag |= 1;
        if (rng_unit(&r) < 0.9) flag |= 2;                   /* proper pair (mostly)     */
        if (rng_unit(&r) < 0.05) flag |= 8;                  /* mate unmapped sometimes  */
        flag |= second ? 128 : 64;
        if (!reverse) flag |= 32;
        tlen = tlen_ok ? (int64_t)n : (int64_t)n + 1 + (int64_t)rng_below(&r, 300);
        if (reverse) tlen = -tlen;
    }
    if (kind_badflag) {
        static const uint32_t bad[5] = { 4, 256, 512, 1024, 2048 };
        flag |= bad[rng_below(&r, 5)];
    }

    /* cigar */
    if (kind_indel && n >= 12) {
        uint32_t a = 3 + (uint32_t)rng_below(&r, n - 8);
        char *q = cigar;
        if (rng_unit(&r) < 0.5) {       /* insertion: query-consuming */
            q = put_u64(q, a); *q++ = 'M'; *q++ = '1'; *q++ = 'I'; q = put_u64(q, n - a - 1); *q++ = 'M';
        } else {                        /* deletion */
            q = put_u64(q, a); *q++ = 'M'; *q++ = '2'; *q++ = 'D'; q = put_u64(q, n - a); *q++ = 'M';
        }
        *q = '\0';
    } else if (kind_clip && n >= 12) {
        uint32_t a = 1 + (uint32_t)rng_below(&r, 8);
        char *q = cigar;
        if (rng_unit(&r) < 0.5) { q = put_u64(q, a); *q++ = 'S'; q = put_u64(q, n - a); *q++ = 'M'; }
        else                    { q = put_u64(q, n - a); *q++ = 'M'; q = put_u64(q, a); *q++ = 'S'; }
        *q = '\0';
    } else {
        char *q = put_u64(cigar, n); *q++ = 'M'; *q = '\0';
    }

    /* record */
    *p++ = 'r'; p = put_u64(p, idx); *p++ = '\t';
    p = put_u64(p, flag); *p++ = '\t';
    if (kind_unk) p = put_str(p, "chrUn_synthetic_decoy"); else p = put_str(p, cname[ci]);
    *p++ = '\t';
    p = put_u64(p, start + 1); *p++ = '\t';
    p = put_u64(p, mapq); *p++ = '\t';
    p = put_str(p, cigar); *p++ = '\t';
    if (kind_paired) { *p++ = '='; *p++ = '\t'; p = put_u64(p, start + 1); }
    else             { *p++ = '*'; *p++ = '\t'; *p++ = '0'; }
    *p++ = '\t';
    if (tlen < 0) { *p++ = '-'; p = put_u64(p, (uint64_t)(-tlen)); } else p = put_u64(p, (uint64_t)tlen);
    *p++ = '\t';
    memcpy(p, seq, n); p += n; *p++ = '\t';
    if (kind_qstar) { *p++ = '*'; }
    else { memset(p, 'I', n); p += n; }
    if (C->with_tags) {
        /* NM + MD (MD is never parsed by the reference; kept for realistic record size) */
        uint32_t run = 0;
        p = put_str(p, "\tNM:i:"); p = put_u64(p, nm);
        p = put_str(p, "\tMD:Z:");
        for (i = 0; i < n; i++) {
            if (seq[i] == refw[i]) run++;
            else { p = put_u64(p, run); *p++ = refw[i]; run = 0; }
        }
        p = put_u64(p, run);
    }
    *p++ = '\n';
    return (size_t)(p - out);
}

static double *cum_weights(const uint64_t *clen, uint32_t nc)
{
    double *c = (double *)malloc(sizeof(double) * nc), tot = 0, run = 0;
    uint32_t i;
    for (i = 0; i < nc; i++) tot += (double)clen[i];
    for (i = 0; i < nc; i++) { run += (double)clen[i]; c[i] = run / tot; }
    return c;
}

/* Upper bound of bytes for reads [begin,end). */
size_t synth_sam_bound(const synth_reads_cfg *C, uint64_t begin, uint64_t end)
{
    size_t per = 2 * (size_t)C->max_len + 160 + (C->with_tags ? (size_t)C->max_len / 2 + 48 : 0);
    return (size_t)(end - begin) * per + 64;
}

/* SAM text of reads [begin,end) into out (capacity out_cap); returns bytes
 * written, or 0 if out_cap is too small.  Deterministic for any thread count. */
size_t synth_sam(const synth_reads_cfg *C,
                 const char *const *cseq, const uint64_t *clen, const char *const *cname,
                 uint32_t nc, uint64_t begin, uint64_t end, char *out, size_t out_cap)
{
    double  *ccum = cum_weights(clen, nc);
    uint64_t n = end - begin;
    uint64_t blk = 4096, nblk = (n + blk - 1) / blk;
    size_t  *bsz = (size_t *)calloc(nblk + 1, sizeof(size_t));
    char   **bbuf = (char **)calloc(nblk, sizeof(char *));
    size_t   total = 0;
    int64_t  b;

#pragma omp parallel for schedule(dynamic, 4)
    for (b = 0; b < (int64_t)nblk; b++) {
        uint64_t i0 = begin + (uint64_t)b * blk, i1 = i0 + blk > end ? end : i0 + blk, i;
        size_t   cap = (size_t)(i1 - i0) * SYNTH_MAX_REC / 8 + SYNTH_MAX_REC, used = 0;
        char    *buf = (char *)malloc(cap);
        for (i = i0; i < i1; i++) {
            if (used + SYNTH_MAX_REC > cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
            used += one_read(C, i, cseq, clen, cname, ccum, nc, buf + used);
        }
        bbuf[b] = buf; bsz[b] = used;
    }
    for (b = 0; b < (int64_t)nblk; b++) total += bsz[b];
    if (total > out_cap) total = 0;
    else {
        size_t *off = (size_t *)malloc(sizeof(size_t) * (nblk + 1));
        off[0] = 0;
        for (b = 0; b < (int64_t)nblk; b++) off[b + 1] = off[b] + bsz[b];
#pragma omp parallel for schedule(static)
        for (b = 0; b < (int64_t)nblk; b++) memcpy(out + off[b], bbuf[b], bsz[b]);
        free(off);
    }
    for (b = 0; b < (int64_t)nblk; b++) free(bbuf[b]);
    free(bbuf); free(bsz); free(ccum);
    return total;
}

/* FASTA text of one contig: ">name\n" + 60-column lines. Returns bytes. */
size_t synth_fasta_record(const char *name, const char *seq, uint64_t len, uint32_t width, char *out)
{
    char    *p = out;
    uint64_t i;
    *p++ = '>'; p = put_str(p, name); *p++ = '\n';
    for (i = 0; i < len; i += width) {
        uint64_t w = i + width > len ? len - i : width;
        memcpy(p, seq + i, w); p += w; *p++ = '\n';
    }
    return (size_t)(p - out);
}
