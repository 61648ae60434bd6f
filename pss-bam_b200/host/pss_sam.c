/* pss_sam.c -- bounded SAM line scanner behind the sam-parse.h API. */
#include "pss_sam.h"

#include <stdint.h>

static int ws(unsigned c) { return c == ' ' || (c - 9u) <= 4u; }   /* isspace(), "C" locale */

/* one %s conversion: skip blanks, copy a run of non-blank bytes (bounded) */
static const char *take_text(const char *p, char *dst, size_t *len_out)
{
    size_t n = 0;
    while (ws((unsigned char)*p)) p++;
    if (!*p) return NULL;
    while (*p && !ws((unsigned char)*p)) {
        if (n == MAX_FIELD_WIDTH) return NULL;         /* the reference would overflow its field here */
        dst[n++] = *p++;
    }
    dst[n] = '\0';
    if (len_out) *len_out = n;
    return p;
}

/* one %u / %lu / %i conversion with strtoul / strtol semantics */
static const char *take_number(const char *p, int is_i, uint64_t *out)
{
    unsigned base = 10;
    int      neg = 0, digits = 0, overflow = 0;
    uint64_t v = 0;
    while (ws((unsigned char)*p)) p++;
    if (!*p) return NULL;
    if (*p == '+' || *p == '-') neg = (*p++ == '-');
    if (is_i && *p == '0') {
        if ((p[1] | 32) == 'x') { base = 16; p += 2; digits = 1; }   /* "0x" alone converts to 0 */
        else base = 8;
    }
    for (;; p++) {
        unsigned c = (unsigned char)*p, d;
        if (c - '0' <= 9u) d = c - '0';
        else if (base == 16 && ((c | 32) - 'a') <= 5u) d = (c | 32) - 'a' + 10;
        else break;
        if (d >= base) break;
        if (v > (UINT64_MAX - d) / base) overflow = 1; else v = v * base + d;
        digits++;
    }
    if (!digits) return NULL;
    if (is_i) {
        int64_t x;
        if (!neg) x = (overflow || v > (uint64_t)INT64_MAX) ? INT64_MAX : (int64_t)v;
        else      x = (overflow || v > (uint64_t)INT64_MAX + 1u) ? INT64_MIN : (int64_t)(0 - v);
        *out = (uint64_t)x;
    } else {
        *out = overflow ? UINT64_MAX : (neg ? 0 - v : v);
    }
    return p;
}

int line2saml(const char *line, Saml *sp)
{
    const char *p = line;
    uint64_t    num;
    size_t      seq_len = 0, qual_len = 0;
    char        mrnm[MAX_FIELD_WIDTH + 1], qname[MAX_FIELD_WIDTH + 1];

    if (!(p = take_text(p, qname, NULL))) return 1;
    if (!(p = take_number(p, 0, &num))) return 1;
    sp->flag = (unsigned int)num;
    if (!(p = take_text(p, sp->rname, NULL))) return 1;
    if (!(p = take_number(p, 0, &num))) return 1;
    sp->pos = (unsigned long)num;
    if (!(p = take_number(p, 0, &num))) return 1;
    sp->mapq = (unsigned int)num;
    if (!(p = take_text(p, sp->cigar, NULL))) return 1;
    if (!(p = take_text(p, mrnm, NULL))) return 1;
    if (!(p = take_number(p, 0, &num))) return 1;
    sp->mpos = (unsigned int)num;
    if (!(p = take_number(p, 1, &num))) return 1;
    sp->isize = (int)(uint32_t)num;
    if (!(p = take_text(p, sp->seq, &seq_len))) return 1;
    if (!(p = take_text(p, sp->qual, &qual_len))) return 1;
    if (seq_len != qual_len) return 1;
    memcpy(sp->qname, qname, sizeof qname);
    memcpy(sp->mrnm, mrnm, sizeof mrnm);

    sp->seq_len = (int)seq_len;
    sp->paired        = (sp->flag >> 0) & 1u;
    sp->proper_pair   = (sp->flag >> 1) & 1u;
    sp->unmap         = (sp->flag >> 2) & 1u;
    sp->munmap        = (sp->flag >> 3) & 1u;
    sp->reverse       = (sp->flag >> 4) & 1u;
    sp->mreverse      = (sp->flag >> 5) & 1u;
    sp->read1         = (sp->flag >> 6) & 1u;
    sp->read2         = (sp->flag >> 7) & 1u;
    sp->secondary     = (sp->flag >> 8) & 1u;
    sp->qc_failed     = (sp->flag >> 9) & 1u;
    sp->duplicate     = (sp->flag >> 10) & 1u;
    sp->supplementary = (sp->flag >> 11) & 1u;
    if (!sp->paired) sp->isize = (int)seq_len;       /* merged / single reads: template length = read length */

    {   /* optional fields: whatever follows the 11th tab, if it fits */
        const char *q = line;
        int tabs = 0;
        while (*q && tabs < 11) tabs += (*q++ == '\t');
        if (*q && strlen(q) <= MAX_FIELD_WIDTH) strcpy(sp->tags, q);
    }
    return 0;
}

int is_header(const char *line) { return line[0] == '@'; }

int aln_seq_len(const char *cigar)
{
    long total = 0, run = 0;
    for (; *cigar; cigar++) {
        if (isdigit((unsigned char)*cigar)) run = run * 10 + (*cigar - '0');
        else { if (*cigar == 'M') total += run; run = 0; }
    }
    return (int)total;
}

int good_score(Saml *sp, float m, float b)
{
    if (sp->AS <= 0) return 1;
    return (float)sp->AS >= m * (float)sp->seq_len + b;
}
