// pss_bamrec.h -- BAM records (SAM spec 4.2) on the device: finding them in the inflated stream and rendering each
// one as the SAM line `samtools view` would hand to the reference's fgets loop (pss-bam.c:148-162,764), reduced to
// what that loop can observe.  Host + device code (the CPU test-suite runs it against a Python BAM reader; the
// product only executes it inside CUDA kernels).
//
// Why text: line2saml (sam-parse.c:10-91) and process_aln (pss-bam.c:390-496, fragkon.c:122-216) are defined on the
// SAM line, glibc sscanf rules included, and the tally kernel reproduces them bit for bit.  Rendering the record and
// feeding that kernel keeps ONE implementation of those semantics; the text costs ~1.2 bytes per inflated byte of HBM
// traffic, far below what the inflate stage costs.  Only what the reference reads is rendered faithfully:
//   QNAME  "q"            (never read: sam-parse.c:37 stores it, nothing uses it; BAM names cannot hold white space)
//   FLAG, RNAME (name of refID in the BAM header, "*" for -1), POS + 1, MAPQ, CIGAR (ops "MIDNSHP=XB", "*" if none),
//   TLEN, SEQ ("=ACMGRSVTWYHKDBN", "*" if empty)                                      -- exactly what samtools prints
//   RNEXT  "*", PNEXT "0" (parsed by the %s / %u conversions, values never read)
//   QUAL   "*" when SEQ is empty or the first quality is 0xff (samtools' rule), else l_seq times 'I': line2saml only
//          compares strlen(qual) with strlen(seq) (sam-parse.c:50); Phred+33 of a valid quality (<= 93) is never white
//          space, so the length is all that is observable
//   tags   dropped (sp->tags is never read)
#pragma once

#include <stdint.h>

#include "pss_record.h"

namespace pssgpu {

constexpr uint32_t kBamMaxRecord = (16u << 20) - 64u;     // block_size beyond this: refused (PSSGPU_EUNSUPP at sync)
constexpr uint32_t kBamFixed     = 32;                    // fixed-length part of a record after block_size

struct BamRefs {                     // reference dictionary of the BAM header, resident on the device
    const uint8_t  *blob;            // the header bytes
    const uint32_t *name_off;        // into blob
    const uint32_t *name_len;        // without the NUL
    int32_t         n_ref;
};

// little-endian u32 at any alignment.  Device: two aligned word loads and a funnel shift (up to 7 bytes behind p are
// touched: every buffer this is used on has that slack).
PSS_HD uint32_t bam_u32(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], 8u * (uint32_t)(a & 3u));
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}

struct BamCore {
    uint32_t block_size;
    int32_t  ref_id, pos;
    uint32_t l_read_name, mapq, n_cigar, flag, l_seq;
    int32_t  next_ref_id, next_pos, tlen;
};
PSS_HD BamCore bam_core(const uint8_t *r)
{
    BamCore c;
    c.block_size = bam_u32(r);
    c.ref_id = (int32_t)bam_u32(r + 4);
    c.pos = (int32_t)bam_u32(r + 8);
    const uint32_t w12 = bam_u32(r + 12), w16 = bam_u32(r + 16);
    c.l_read_name = w12 & 0xffu;
    c.mapq = (w12 >> 8) & 0xffu;
    c.n_cigar = w16 & 0xffffu;
    c.flag = w16 >> 16;
    c.l_seq = bam_u32(r + 20);
    c.next_ref_id = (int32_t)bam_u32(r + 24);
    c.next_pos = (int32_t)bam_u32(r + 28);
    c.tlen = (int32_t)bam_u32(r + 32);
    return c;
}
// bytes of the record's variable part that SEQ/QUAL/CIGAR/name claim; <= block_size - 32 in a well-formed record
PSS_HD uint64_t bam_claimed(const BamCore &c)
{
    return (uint64_t)c.l_read_name + 4ull * c.n_cigar + (((uint64_t)c.l_seq + 1) >> 1) + (uint64_t)c.l_seq;
}
PSS_HD bool bam_wellformed(const BamCore &c)
{
    return c.block_size >= kBamFixed && c.block_size <= kBamMaxRecord && c.l_read_name >= 1u && c.l_seq <= 0x7fffffffu
        && bam_claimed(c) <= (uint64_t)(c.block_size - kBamFixed);
}

// Could a record start at u + x?  Used only to GUESS where the first record of a BGZF block starts (records do not
// align with blocks and carry no sync marks); every guess is verified against the exact chain of the preceding block,
// so a false positive costs time, never correctness.  `end` = end of the inflated data at hand.
PSS_HD bool bam_plausible_at(const uint8_t *u, uint64_t x, uint64_t end, int32_t n_ref)
{
    if (x + 4 + kBamFixed > end) return false;
    const BamCore c = bam_core(u + x);
    if (!bam_wellformed(c)) return false;
    if (c.ref_id < -1 || c.ref_id >= n_ref || c.next_ref_id < -1 || c.next_ref_id >= n_ref) return false;
    if (c.pos < -1 || c.next_pos < -1) return false;
    const uint64_t nul = x + 4 + kBamFixed + c.l_read_name - 1;
    if (nul < end && u[nul] != 0) return false;
    return true;
}
// three records in a row (as far as the data reaches)
PSS_HD bool bam_guess_at(const uint8_t *u, uint64_t x, uint64_t end, int32_t n_ref)
{
    for (int k = 0; k < 3; k++) {
        if (!bam_plausible_at(u, x, end, n_ref)) return k > 0 && x + 4 + kBamFixed > end;   // ran out of data after >= 1 good record
        x += 4ull + bam_u32(u + x);
    }
    return true;
}

// ---- RG filter (samtools view -r RG: keep records whose RG:Z tag equals the name; records without one are dropped)
PSS_HD bool bam_has_read_group(const uint8_t *r, const BamCore &c, const char *rg, int rg_len)
{
    const uint8_t *p = r + 4 + kBamFixed + bam_claimed(c), *e = r + 4 + c.block_size;
    while (p + 3 <= e) {
        const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
        p += 3;
        if (ty == 'Z' || ty == 'H') {
            const uint8_t *v = p;
            while (p < e && *p) p++;
            if (t0 == 'R' && t1 == 'G' && ty == 'Z') {        // the first RG tag decides, as bam_aux_get does
                if ((int)(p - v) != rg_len) return false;
                for (int i = 0; i < rg_len; i++) if (v[i] != (uint8_t)rg[i]) return false;
                return true;
            }
            p++;                                             // the NUL
        } else if (ty == 'A' || ty == 'c' || ty == 'C') p += 1;
        else if (ty == 's' || ty == 'S') p += 2;
        else if (ty == 'i' || ty == 'I' || ty == 'f') p += 4;
        else if (ty == 'B') {
            if (p + 5 > e) return false;
            const uint8_t  st = p[0];
            const uint32_t n = bam_u32(p + 1);
            const uint32_t sz = (st == 'c' || st == 'C') ? 1u : (st == 's' || st == 'S') ? 2u : 4u;
            if ((uint64_t)n * sz > (uint64_t)(e - p)) return false;
            p += 5 + (uint64_t)n * sz;
        } else return false;                                 // unknown type: the rest cannot be walked
    }
    return false;
}

// ---- rendering ---------------------------------------------------------------------------------------------------------
struct BamCountSink {
    uint32_t n = 0;
    PSS_HD void put(uint8_t) { n++; }
    PSS_HD void fill(uint8_t, uint32_t k) { n += k; }
};
struct BamWriteSink {
    uint8_t *p;
    PSS_HD void put(uint8_t c) { *p++ = c; }
    PSS_HD void fill(uint8_t c, uint32_t k) { for (uint32_t i = 0; i < k; i++) p[i] = c; p += k; }
};
template <class Sink>
PSS_HD void bam_put_u32(Sink &s, uint32_t v)
{
    uint8_t d[10];
    int     n = 0;
    do { d[n++] = (uint8_t)('0' + v % 10u); v /= 10u; } while (v);
    while (n) s.put(d[--n]);
}
template <class Sink>
PSS_HD void bam_put_i64(Sink &s, int64_t v)
{
    if (v < 0) { s.put('-'); bam_put_u32(s, (uint32_t)(0 - v)); }     // |v| <= 2^31 here
    else bam_put_u32(s, (uint32_t)v);
}

// One record -> one line, in two parts (the render kernel writes the second one with all lanes of a warp).
// r points at block_size.  c must be bam_wellformed().
//   head: QNAME .. TLEN and the tab behind it        tail: SEQ, tab, QUAL, newline
template <class Sink>
PSS_HD void bam_render_head(const uint8_t *r, const BamCore &c, const BamRefs &R, Sink &s)
{
    const uint8_t *cig = r + 4 + kBamFixed + c.l_read_name;
    s.put('q'); s.put('\t');
    bam_put_u32(s, c.flag); s.put('\t');
    if (c.ref_id >= 0 && c.ref_id < R.n_ref) {
        const uint8_t *nm = R.blob + R.name_off[c.ref_id];
        const uint32_t nl = R.name_len[c.ref_id];
        if (nl == 0) s.put('*');                               // an empty name is not valid BAM; keep the field count
        for (uint32_t i = 0; i < nl; i++) s.put(nm[i]);
    } else s.put('*');
    s.put('\t');
    bam_put_i64(s, (int64_t)c.pos + 1); s.put('\t');
    bam_put_u32(s, c.mapq); s.put('\t');
    if (c.n_cigar == 0) s.put('*');
    for (uint32_t i = 0; i < c.n_cigar; i++) {
        const uint32_t v = bam_u32(cig + 4 * i);
        bam_put_u32(s, v >> 4);
        const uint32_t op = v & 15u;
        // "MIDNSHP=XB", '?' beyond (htslib's BAM_CIGAR_STR padding)
        s.put(op < 10u ? (uint8_t)"MIDNSHP=XB"[op] : (uint8_t)'?');
    }
    s.put('\t'); s.put('*'); s.put('\t'); s.put('0'); s.put('\t');
    bam_put_i64(s, (int64_t)c.tlen); s.put('\t');
}
PSS_HD const uint8_t *bam_seq_ptr(const uint8_t *r, const BamCore &c) { return r + 4 + kBamFixed + c.l_read_name + 4ull * c.n_cigar; }
// byte k of the tail (0 <= k < bam_tail_len): what the cooperative writer of the render kernel stores, lane by lane
PSS_HD uint32_t bam_tail_len(const BamCore &c, bool qual_star) { return c.l_seq == 0 ? 4u : c.l_seq + 1u + (qual_star ? 1u : c.l_seq) + 1u; }
PSS_HD bool bam_qual_is_star(const uint8_t *r, const BamCore &c) { return c.l_seq == 0 || bam_seq_ptr(r, c)[(c.l_seq + 1u) >> 1] == 0xffu; }
PSS_HD uint8_t bam_tail_byte(const uint8_t *seq, const BamCore &c, bool qual_star, uint32_t k)
{
    if (c.l_seq == 0) return (uint8_t)"*\t*\n"[k];
    if (k < c.l_seq) {
        const uint32_t b = seq[k >> 1];
        return (uint8_t)"=ACMGRSVTWYHKDBN"[(k & 1u) ? (b & 15u) : (b >> 4)];
    }
    if (k == c.l_seq) return (uint8_t)'\t';
    const uint32_t q = k - c.l_seq - 1u, ql = qual_star ? 1u : c.l_seq;
    if (q < ql) return qual_star ? (uint8_t)'*' : (uint8_t)'I';
    return (uint8_t)'\n';
}
template <class Sink>
PSS_HD void bam_render(const uint8_t *r, const BamCore &c, const BamRefs &R, Sink &s)
{
    bam_render_head(r, c, R, s);
    const uint8_t *seq = bam_seq_ptr(r, c);
    const bool     qs = bam_qual_is_star(r, c);
    const uint32_t n = bam_tail_len(c, qs);
    for (uint32_t k = 0; k < n; k++) s.put(bam_tail_byte(seq, c, qs, k));
}

}  // namespace pssgpu
