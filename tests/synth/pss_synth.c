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

/* ========================================================================== */
/* SAM text -> BAM (BGZF) writer: test + bench infrastructure for the device   */
/* BGZF/BAM ingest.  Follows the SAM specification sections 4.1 (BGZF) and 4.2 */
/* (BAM); blocks hold 0xff00 bytes of payload like htslib's, records run       */
/* across block boundaries, an empty EOF block ends the file.                  */
/* Only well-formed SAM lines (what synth_sam writes) are accepted.            */
/* ========================================================================== */
#include <zlib.h>

typedef struct { uint8_t *p; size_t n, cap; } bytebuf;
static void bb_reserve(bytebuf *b, size_t extra)
{
    if (b->n + extra > b->cap) {
        size_t c = b->cap ? b->cap : 4096;
        while (c < b->n + extra) c *= 2;
        b->p = (uint8_t *)realloc(b->p, c);
        b->cap = c;
    }
}
static void bb_put(bytebuf *b, const void *src, size_t n) { bb_reserve(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static void bb_u32(bytebuf *b, uint32_t v) { uint8_t t[4] = { (uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24) }; bb_put(b, t, 4); }
static void bb_u16(bytebuf *b, uint32_t v) { uint8_t t[2] = { (uint8_t)v, (uint8_t)(v >> 8) }; bb_put(b, t, 2); }
static void bb_u8(bytebuf *b, uint32_t v) { uint8_t t = (uint8_t)v; bb_put(b, &t, 1); }

static int reg2bin(int64_t beg, int64_t end)     /* SAM spec 5.3 */
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
static int ref_index(const char *name, size_t len, const char *const *names, uint32_t n_ref)
{
    uint32_t i;
    for (i = 0; i < n_ref; i++)
        if (strlen(names[i]) == len && memcmp(names[i], name, len) == 0) return (int)i;
    return -1;
}
static int base_code(char c)
{
    const char *tab = "=ACMGRSVTWYHKDBN", *p = strchr(tab, c >= 'a' && c <= 'z' ? c - 32 : c);
    return (p && *p) ? (int)(p - tab) : 15;
}

/* one SAM line [s, e) (no newline) -> BAM record appended to b.  Returns 0, or -1 for a line this writer cannot
 * represent.  idx = index of the line (drives the optional RG tag and quality substitution). */
static int sam_line_to_bam(const char *s, const char *e, uint64_t idx, const char *const *names, uint32_t n_ref,
                           int rg_mode, int qual_mode, bytebuf *b)
{
    const char *f[12];
    size_t      fl[12];
    int         nf = 0;
    const char *p = s, *tags;
    while (nf < 11) {
        const char *q = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!q) { if (nf == 10) { q = e; } else return -1; }
        f[nf] = p; fl[nf] = (size_t)(q - p); nf++;
        p = q < e ? q + 1 : e;
        if (q == e) break;
    }
    if (nf < 11) return -1;
    tags = p;                                       /* after the tab behind QUAL, or e */
    {
        char     tmp[32];
        uint32_t flag, mapq, n_cig = 0, l_seq, cig[256];
        int64_t  pos, pnext, tlen, ref_len = 0;
        int      ref_id, next_id;
        size_t   start, i;
#define NUMF(k, dst) do { if (fl[k] >= sizeof tmp) return -1; memcpy(tmp, f[k], fl[k]); tmp[fl[k]] = 0; dst = strtoll(tmp, NULL, 10); } while (0)
        { int64_t v; NUMF(1, v); flag = (uint32_t)v; NUMF(3, pos); NUMF(4, v); mapq = (uint32_t)v; NUMF(7, pnext); NUMF(8, tlen); }
        ref_id = (fl[2] == 1 && f[2][0] == '*') ? -1 : ref_index(f[2], fl[2], names, n_ref);
        if (ref_id < 0 && !(fl[2] == 1 && f[2][0] == '*')) return -1;
        next_id = (fl[6] == 1 && f[6][0] == '=') ? ref_id : (fl[6] == 1 && f[6][0] == '*') ? -1 : ref_index(f[6], fl[6], names, n_ref);
        if (!(fl[5] == 1 && f[5][0] == '*')) {
            const char *c = f[5], *ce = f[5] + fl[5];
            while (c < ce) {
                uint32_t    len = 0;
                const char *ops = "MIDNSHP=X", *o;
                while (c < ce && *c >= '0' && *c <= '9') len = len * 10 + (uint32_t)(*c++ - '0');
                if (c >= ce || !(o = strchr(ops, *c)) || n_cig >= 256) return -1;
                cig[n_cig++] = (len << 4) | (uint32_t)(o - ops);
                if (*c == 'M' || *c == 'D' || *c == 'N' || *c == '=' || *c == 'X') ref_len += len;
                c++;
            }
        }
        l_seq = (fl[9] == 1 && f[9][0] == '*') ? 0u : (uint32_t)fl[9];
        start = b->n;
        bb_u32(b, 0);                               /* block_size, patched below */
        bb_u32(b, (uint32_t)ref_id);
        bb_u32(b, (uint32_t)(pos - 1));
        bb_u8(b, (uint32_t)fl[0] + 1);
        bb_u8(b, mapq);
        bb_u16(b, (uint32_t)reg2bin(pos - 1, pos - 1 + (ref_len ? ref_len : 1)));
        bb_u16(b, n_cig);
        bb_u16(b, flag);
        bb_u32(b, l_seq);
        bb_u32(b, (uint32_t)next_id);
        bb_u32(b, (uint32_t)(pnext - 1));
        bb_u32(b, (uint32_t)tlen);
        bb_put(b, f[0], fl[0]); bb_u8(b, 0);
        for (i = 0; i < n_cig; i++) bb_u32(b, cig[i]);
        for (i = 0; i < l_seq; i += 2) {
            const int hi = base_code(f[9][i]), lo = i + 1 < l_seq ? base_code(f[9][i + 1]) : 0;
            bb_u8(b, (uint32_t)(hi << 4 | lo));
        }
        if (fl[10] == 1 && f[10][0] == '*' && l_seq != 1) { for (i = 0; i < l_seq; i++) bb_u8(b, 0xff); }
        else if (qual_mode == 1) {
            /* binned, run-structured qualities (4 levels, as recent instruments write): realistic entropy for the
             * compressed size; the tallies never look at them */
            static const uint8_t lev[4] = { 37, 37, 23, 11 };
            rng_t r = rng_at(0x51a7, 11, idx);
            uint32_t run = 0, q = 37;
            for (i = 0; i < l_seq; i++) {
                if (run == 0) { const uint64_t v = rng_u64(&r); q = lev[v & 3]; run = 1 + (uint32_t)((v >> 8) % 12); }
                bb_u8(b, q); run--;
            }
        } else {
            for (i = 0; i < l_seq; i++) bb_u8(b, (uint32_t)((unsigned char)f[10][i] - 33));
        }
        /* optional fields: TAG:i:int and TAG:Z:string */
        while (tags < e) {
            const char *q = (const char *)memchr(tags, '\t', (size_t)(e - tags));
            const char *te = q ? q : e;
            if (te - tags < 5 || tags[2] != ':' || tags[4] != ':') return -1;
            bb_put(b, tags, 2);
            if (tags[3] == 'i') {
                int64_t v;
                size_t  n = (size_t)(te - tags - 5);
                if (n >= sizeof tmp) return -1;
                memcpy(tmp, tags + 5, n); tmp[n] = 0; v = strtoll(tmp, NULL, 10);
                if (v >= 0 && v < 256) { bb_u8(b, 'C'); bb_u8(b, (uint32_t)v); }
                else if (v >= 0 && v < 65536) { bb_u8(b, 'S'); bb_u16(b, (uint32_t)v); }
                else if (v >= 0) { bb_u8(b, 'I'); bb_u32(b, (uint32_t)v); }
                else if (v >= -128) { bb_u8(b, 'c'); bb_u8(b, (uint32_t)v); }
                else if (v >= -32768) { bb_u8(b, 's'); bb_u16(b, (uint32_t)v); }
                else { bb_u8(b, 'i'); bb_u32(b, (uint32_t)v); }
            } else if (tags[3] == 'Z') {
                bb_u8(b, 'Z'); bb_put(b, tags + 5, (size_t)(te - tags - 5)); bb_u8(b, 0);
            } else return -1;
            tags = q ? q + 1 : e;
        }
        if (rg_mode == 1 && idx % 3 != 2) {
            /* every third record has no RG; an array tag in front of it exercises the tag walk */
            static const uint8_t arr[] = { 'X', 'B', 'B', 'S', 2, 0, 0, 0, 7, 0, 9, 0 };
            bb_put(b, arr, sizeof arr);
            bb_put(b, "RGZ", 3);
            bb_put(b, idx % 3 == 0 ? "rgA" : "rgB", 4);
        }
        {
            const uint32_t bs = (uint32_t)(b->n - start - 4);
            b->p[start] = (uint8_t)bs; b->p[start + 1] = (uint8_t)(bs >> 8); b->p[start + 2] = (uint8_t)(bs >> 16); b->p[start + 3] = (uint8_t)(bs >> 24);
        }
#undef NUMF
    }
    return 0;
}

/* BGZF block around payload[0..n): returns bytes written to out (<= 65536 + 26) */
static size_t bgzf_block(const uint8_t *payload, size_t n, int level, uint8_t *out)
{
    static const uint8_t head[12] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0 };
    z_stream zs;
    size_t   clen;
    uint32_t crc = (uint32_t)crc32(crc32(0L, NULL, 0), payload, (uInt)n), total;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef *)payload; zs.avail_in = (uInt)n;
    zs.next_out = out + 18; zs.avail_out = 65536 + 1024;
    deflate(&zs, Z_FINISH);
    clen = zs.total_out;
    deflateEnd(&zs);
    memcpy(out, head, 12);
    out[12] = 'B'; out[13] = 'C'; out[14] = 2; out[15] = 0;
    total = (uint32_t)(18 + clen + 8);
    out[16] = (uint8_t)(total - 1); out[17] = (uint8_t)((total - 1) >> 8);
    out[18 + clen + 0] = (uint8_t)crc; out[18 + clen + 1] = (uint8_t)(crc >> 8); out[18 + clen + 2] = (uint8_t)(crc >> 16); out[18 + clen + 3] = (uint8_t)(crc >> 24);
    out[18 + clen + 4] = (uint8_t)n; out[18 + clen + 5] = (uint8_t)(n >> 8); out[18 + clen + 6] = (uint8_t)(n >> 16); out[18 + clen + 7] = (uint8_t)(n >> 24);
    return total;
}

/* Upper bound of the BAM size for `sam_len` bytes of SAM text. */
size_t synth_bam_bound(size_t sam_len, uint32_t n_ref) { return sam_len + sam_len / 2 + (size_t)n_ref * 600 + (1u << 20); }

/* SAM text (whole lines) -> BAM file bytes.  names/lens: the @SQ dictionary (n_ref entries).  level: zlib level
 * (0 = stored blocks).  rg_mode 1: RG:Z:rgA / rgB / none by line index.  qual_mode 1: binned synthetic qualities.
 * block_payload: bytes per BGZF block (0 = htslib's 0xff00).  Returns the BAM size, 0 on failure (bad line / capacity). */
size_t synth_sam_to_bam(const char *sam, size_t sam_len, const char *const *names, const uint64_t *lens, uint32_t n_ref,
                        int level, int rg_mode, int qual_mode, uint32_t block_payload, uint8_t *out, size_t out_cap)
{
    const size_t chunk = 4u << 20;                  /* SAM bytes per conversion task, cut at newlines */
    size_t   n_tasks = (sam_len + chunk - 1) / chunk, t;
    size_t  *cut = (size_t *)malloc(sizeof(size_t) * (n_tasks + 2));
    uint64_t *first_idx = (uint64_t *)calloc(n_tasks + 2, sizeof(uint64_t));
    bytebuf *parts = (bytebuf *)calloc(n_tasks + 1, sizeof(bytebuf));
    bytebuf  hdr = { NULL, 0, 0 }, all = { NULL, 0, 0 };
    int      bad = 0;
    size_t   total = 0, n_blocks, bi, out_n = 0;
    uint32_t i;
    if (block_payload == 0 || block_payload > 0xff00) block_payload = 0xff00;
    cut[0] = 0;
    for (t = 1; t <= n_tasks; t++) {
        size_t at = t * chunk;
        if (at >= sam_len) at = sam_len;
        else { const char *q = (const char *)memchr(sam + at, '\n', sam_len - at); at = q ? (size_t)(q - sam) + 1 : sam_len; }
        if (at < cut[t - 1]) at = cut[t - 1];
        cut[t] = at;
    }
    /* line index of the first line of every task */
#pragma omp parallel for schedule(dynamic, 1)
    for (t = 0; t < n_tasks; t++) {
        uint64_t c = 0; size_t k;
        for (k = cut[t]; k < cut[t + 1]; k++) c += sam[k] == '\n';
        first_idx[t + 1] = c;
    }
    for (t = 1; t <= n_tasks; t++) first_idx[t] += first_idx[t - 1];
#pragma omp parallel for schedule(dynamic, 1)
    for (t = 0; t < n_tasks; t++) {
        const char *p = sam + cut[t], *e = sam + cut[t + 1];
        uint64_t    idx = first_idx[t];
        while (p < e) {
            const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *le = q ? q : e;
            if (le > p && sam_line_to_bam(p, le, idx, names, n_ref, rg_mode, qual_mode, &parts[t]) != 0) {
#pragma omp atomic write
                bad = 1;
            }
            idx++;
            p = q ? q + 1 : e;
        }
    }
    /* header */
    {
        bytebuf text = { NULL, 0, 0 };
        char    line[700];
        bb_put(&text, "@HD\tVN:1.6\tSO:unsorted\n", 23);
        for (i = 0; i < n_ref; i++) {
            int n = snprintf(line, sizeof line, "@SQ\tSN:%s\tLN:%llu\n", names[i], (unsigned long long)lens[i]);
            bb_put(&text, line, (size_t)n);
        }
        bb_put(&hdr, "BAM\1", 4);
        bb_u32(&hdr, (uint32_t)text.n);
        bb_put(&hdr, text.p, text.n);
        bb_u32(&hdr, n_ref);
        for (i = 0; i < n_ref; i++) {
            const size_t nl = strlen(names[i]) + 1;
            bb_u32(&hdr, (uint32_t)nl); bb_put(&hdr, names[i], nl); bb_u32(&hdr, (uint32_t)lens[i]);
        }
        free(text.p);
    }
    total = hdr.n;
    for (t = 0; t < n_tasks; t++) total += parts[t].n;
    all.p = (uint8_t *)malloc(total ? total : 1); all.cap = total; all.n = 0;
    memcpy(all.p, hdr.p, hdr.n); all.n = hdr.n;
    for (t = 0; t < n_tasks; t++) { memcpy(all.p + all.n, parts[t].p, parts[t].n); all.n += parts[t].n; free(parts[t].p); }
    free(parts); free(hdr.p); free(cut); free(first_idx);
    if (bad) { free(all.p); return 0; }
    /* BGZF: compress the blocks in parallel, then lay them out */
    n_blocks = (all.n + block_payload - 1) / block_payload;
    {
        uint8_t **cb = (uint8_t **)calloc(n_blocks + 1, sizeof(uint8_t *));
        size_t   *cl = (size_t *)calloc(n_blocks + 1, sizeof(size_t));
#pragma omp parallel for schedule(dynamic, 8)
        for (bi = 0; bi < n_blocks; bi++) {
            const size_t o = bi * block_payload, n = o + block_payload > all.n ? all.n - o : block_payload;
            cb[bi] = (uint8_t *)malloc(65536 + 2048);
            cl[bi] = bgzf_block(all.p + o, n, level, cb[bi]);
        }
        for (bi = 0; bi < n_blocks; bi++) {
            if (out_n + cl[bi] + 28 > out_cap) { out_n = 0; bad = 1; break; }
            memcpy(out + out_n, cb[bi], cl[bi]); out_n += cl[bi];
        }
        for (bi = 0; bi < n_blocks; bi++) free(cb[bi]);
        free(cb); free(cl);
    }
    if (!bad) {
        static const uint8_t eof[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        memcpy(out + out_n, eof, 28); out_n += 28;
    }
    free(all.p);
    return bad ? 0 : out_n;
}
