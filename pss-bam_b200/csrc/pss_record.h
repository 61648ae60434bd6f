// pss_record.h -- per-record logic of the pss-bam / fragkon hot path, written
// once for both the sm_100a kernels (pss_kernels.cuh) and a host build that
// the CPU test-suite fuzzes against the oracle (tests/host_emul).  The host
// build is a TEST of this logic; the product library (libpssgpu.so) only ever
// executes it inside CUDA kernels.
//
// What lives here:
//   * the packed-genome layout (DevGenome) and its 16-symbol alphabet
//   * scan11(): an exact emulation of the reference's 11-conversion sscanf
//     (sam-parse.c:36-48) for lines that are not clean tab-separated SAM
//   * split_fast(): the common case -- clean records located through the
//     separator bit masks the tile scan produced
//   * pss_record(): filters + window gather of pss-bam.c:390-496 expressed on
//     2-bit streams (no string copies, no reverse-complement pass)
//   * fk_record(): fragkon.c:122-216
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PSS_HD __host__ __device__ __forceinline__
#define PSS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PSS_HD inline
#define PSS_HD_NOINLINE inline
#endif

namespace pssgpu {

// ---------------------------------------------------------------------------
// Packed genome.
//
// 16 bases per 64-bit "group": low 32 bits = 16 x 2-bit base code (base j of
// the group at bits 2j..2j+1), high 32 bits = 16 x 2-bit class.  The pair
// (class, code) names one of 16 symbols:
//   class 0: A C G T      (code = the reference's A=0 C=1 G=2 T=3, kmer.c:196-211)
//   class 1: N R Y S
//   class 2: W K M B
//   class 3: D H V other  ("other" = any byte not listed; its exact value is
//                           kept in a sorted exception list, needed only for
//                           the -U/-D membership test, pss-bam.c:134-142)
// Keeping code and class of a base in the same 8-byte word means one window
// gather (R+2 <= 32 bases) touches 16-24 contiguous bytes = one 32-byte
// sector in most cases, instead of one sector in a 2-bit plane plus one in a
// mask plane.
//
// Every contig starts on a group boundary and is surrounded by >= kPadBases of
// "other" symbols, so windows that hang over a contig end (fragkon.c:156-177
// can ask for them) read invalid symbols instead of a neighbouring contig.
// ---------------------------------------------------------------------------
constexpr int      kBasesPerGroup = 16;
constexpr int      kPadGroups     = 4;                       // 64 bases each side
constexpr int      kPadBases      = kPadGroups * kBasesPerGroup;
constexpr uint32_t kSymOther      = 15;

struct DevContig {
    uint64_t base_off;    // global base index of position 0 (multiple of 16)
    uint64_t len;
    uint32_t name_off;    // into DevGenome::names
    uint32_t name_len;
};

struct DevGenome {
    const uint64_t  *groups;
    uint64_t         n_groups;
    const DevContig *contigs;
    uint32_t         n_contigs;
    const char      *names;
    const uint32_t  *hash;        // open addressing; value = contig index + 1, 0 = empty
    uint32_t         hash_mask;
    const uint64_t  *exc_pos;     // sorted global base indices of "other" symbols
    const uint8_t   *exc_chr;
    uint32_t         n_exc;
};

constexpr uint32_t kNameHashSeed = 2166136261u;
PSS_HD uint32_t name_hash_step(uint32_t h, uint8_t c) { return (h ^ c) * 16777619u; }

#define PSSGPU_SYM_CHARS "ACGTNRYSWKMBDHV"

// Symbol (0..15) of an upper-cased byte.
PSS_HD uint32_t sym_of_upper(uint8_t c)
{
    switch (c) {
    case 'A': return 0;  case 'C': return 1;  case 'G': return 2;  case 'T': return 3;
    case 'N': return 4;  case 'R': return 5;  case 'Y': return 6;  case 'S': return 7;
    case 'W': return 8;  case 'K': return 9;  case 'M': return 10; case 'B': return 11;
    case 'D': return 12; case 'H': return 13; case 'V': return 14;
    default:  return kSymOther;
    }
}

// toupper() in the C locale (fasta-genome-io.c:127, pss-bam.c:84-89)
PSS_HD uint8_t upper_c(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// One packed group from 16 upper-cased-or-not ASCII bytes; `n_valid` bases are
// real, the rest of the group is padding ("other").
template <class ByteAt>
PSS_HD uint64_t pack_group(const ByteAt &at, int n_valid, uint32_t *other_mask, uint32_t *nul_seen)
{
    uint32_t codes = 0, classes = 0, om = 0;
    for (int j = 0; j < 16; j++) {
        uint32_t s = kSymOther;
        if (j < n_valid) {
            uint8_t c = at(j);
            if (c == 0) *nul_seen = 1;
            s = sym_of_upper(upper_c(c));
            if (s == kSymOther) om |= 1u << j;
        }
        codes   |= (s & 3u) << (2 * j);
        classes |= (s >> 2) << (2 * j);
    }
    *other_mask = om;
    return (uint64_t)codes | ((uint64_t)classes << 32);
}

// ---------------------------------------------------------------------------
// small bit helpers with host twins
// ---------------------------------------------------------------------------
PSS_HD int ffs32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
PSS_HD uint32_t brev32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
// low 32 bits of (hi:lo) >> sh, 0 <= sh <= 31
PSS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
// low 32 bits of (hi:lo) >> sh, 0 <= sh <= 32 (clamping variant: a shift by 32 returns hi)
PSS_HD uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(lo, hi, sh);
#else
    return sh >= 32 ? hi : (sh ? (lo >> sh) | (hi << (32 - sh)) : lo);
#endif
}
// reverse the order of the sixteen 2-bit fields of x
PSS_HD uint32_t rev_fields32(uint32_t x)
{
    uint32_t y = brev32(x);
    return ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
}
PSS_HD uint64_t low_fields_mask(int n_fields)   // n_fields in [0,32]
{
    return n_fields >= 32 ? ~0ull : ((1ull << (2 * n_fields)) - 1ull);
}
// reverse the order of the first n (1..32) 2-bit fields of x; higher fields must be 0
PSS_HD uint64_t rev_fields64(uint64_t x, int n)
{
    uint64_t r = ((uint64_t)rev_fields32((uint32_t)x) << 32) | rev_fields32((uint32_t)(x >> 32));
    return r >> (64 - 2 * n);
}

// ---------------------------------------------------------------------------
// outcomes and options
// ---------------------------------------------------------------------------
enum : int { kModePss = 0, kModeFragkon = 1, kModeBoth = 2 };   // kModeBoth: both tallies from one scan of the text

enum : int {          // per-record outcomes (pssgpu.h pssgpu_debug_fetch)
    kCounted = 0, kNoContig = 1, kFiltered = -1, kParseFail = -2, kUndefined = -3,
    kNeedSlow = 100   // internal: split_fast() declined, run scan11()
};
enum : int { kStLines = 0, kStCounted, kStNoContig, kStFiltered, kStParseFail, kStUndefined, kStN };

constexpr int kMaxCtxChars = 48;
constexpr int kMaxRegion   = 30;        // R + 2 <= 32 bases per window: the ballot tally of the kernels
constexpr int kMaxRegionWide = 2045;    // larger -r go through pss_record_wide(); a counted read has n <= 2047 bases
constexpr int kMaxFragK    = 14;
constexpr int kMaxField    = 2047;      // sam-parse.h:10 MAX_FIELD_WIDTH - 1
constexpr int kMaxLine     = 200000;    // sam-parse.h:8  MAX_LINE_LEN (fgets chunk, pss-bam.c:764)
constexpr int kMaxTlen     = 1000000;   // beyond this the reference's VLAs overflow its stack (pss-bam.c:402)

struct TallyCfg {
    int      mode;
    // pss-bam.c:12-18 / fragkon.c:14-18
    int      R;
    uint64_t min_len, max_len;
    uint32_t min_mq;                    // compared unsigned, like `sp->mapq < MIN_MQ`
    uint32_t merged_only;
    uint32_t up_mask, down_mask;        // bit s set: symbol s (0..14) is in -U / -D
    uint32_t up_other, down_other;      // ctx string holds bytes outside the 15 named symbols
    char     up_ctx[kMaxCtxChars], down_ctx[kMaxCtxChars];
    int      K;
    uint32_t zero;                      // 0, but a run-time value: see fetch_window()
};

// one SAM line after the 11 conversions of sam-parse.c:36-48 (offsets are
// relative to whatever byte accessor produced it)
struct RecView {
    uint32_t flag;
    uint64_t pos;
    uint32_t mapq;
    int32_t  tlen;
    int32_t  rname_off, rname_len;
    int32_t  cigar_off, cigar_len;
    int32_t  seq_off, seq_len;
};

PSS_HD bool is_ws(uint32_t c) { return c == 32u || (c - 9u) <= 4u; }     // isspace(), C locale
PSS_HD bool is_digit(uint32_t c) { return (c - 48u) <= 9u; }

// ---------------------------------------------------------------------------
// scan11: the reference's
//   sscanf(line, "%s\t%u\t%s\t%lu\t%u\t%s\t%s\t%u\t%i\t%s\t%s", ...)   sam-parse.c:36-48
// on bytes b(0..L), following glibc's conversion rules: every conversion and
// every "\t" directive skips any run of isspace() bytes; %s takes a maximal
// run of non-space bytes; %u / %lu take [+-]digits and stop at the first other
// byte (the rest of the token is left for the next conversion) with strtoul's
// overflow rule; %i also takes 0x... and 0... with strtol's clamping; a NUL
// byte ends the string.
//   return kCounted (0)  -> 11 conversions succeeded and strlen(seq)==strlen(qual)
//          kParseFail     -> line2saml returns 1
//          kUndefined     -> one of the first eleven white-space delimited
//                            runs is longer than the reference's 2048-byte
//                            fields: its sscanf would overflow them.
// ---------------------------------------------------------------------------
template <class B>
PSS_HD_NOINLINE int scan11(const B &b, int L, RecView &r)
{
    {
        int i = 0, runs = 0;
        while (i < L && runs < 11) {
            uint32_t c = b(i);
            if (c == 0) break;
            if (is_ws(c)) { i++; continue; }
            int st = i;
            while (i < L) { c = b(i); if (c == 0 || is_ws(c)) break; i++; }
            if (i - st > kMaxField) return kUndefined;
            runs++;
        }
    }
    int i = 0;
    int qual_len = 0;
    for (int conv = 0; conv < 11; conv++) {
        uint32_t c;
        for (;;) {                               // leading white space
            c = i < L ? b(i) : 0u;
            if (!is_ws(c)) break;
            i++;
        }
        if (c == 0) return kParseFail;           // input failure: fewer than 11 conversions
        const bool is_str = (conv == 0 || conv == 2 || conv == 5 || conv == 6 || conv == 9 || conv == 10);
        if (is_str) {
            int st = i;
            while (i < L) { c = b(i); if (c == 0 || is_ws(c)) break; i++; }
            int len = i - st;
            if (conv == 2)  { r.rname_off = st; r.rname_len = len; }
            if (conv == 5)  { r.cigar_off = st; r.cigar_len = len; }
            if (conv == 9)  { r.seq_off = st;   r.seq_len = len; }
            if (conv == 10) qual_len = len;
        } else {
            bool neg = false;
            if (c == '+' || c == '-') { neg = (c == '-'); i++; c = i < L ? b(i) : 0u; }
            uint32_t base = 10;
            int      nd = 0;
            uint64_t v = 0;
            bool     ovf = false;
            if (conv == 8 && c == '0') {         // %i: prefix decides the base
                uint32_t c1 = (i + 1 < L) ? b(i + 1) : 0u;
                if ((c1 | 32u) == 'x') { base = 16; i += 2; nd = 1; }   // "0x" alone still converts (to 0)
                else base = 8;
            }
            for (;;) {
                c = i < L ? b(i) : 0u;
                uint32_t d;
                if (is_digit(c)) d = c - '0';
                else if (base == 16 && (((c | 32u) - 'a') <= 5u)) d = (c | 32u) - 'a' + 10;
                else break;
                if (d >= base) break;
                if (v > (~0ull - d) / base) ovf = true; else v = v * base + d;
                nd++;
                i++;
            }
            if (nd == 0) return kParseFail;      // matching failure
            if (conv == 8) {                     // strtol
                int64_t x;
                if (!neg) x = (ovf || v > 0x7fffffffffffffffull) ? 0x7fffffffffffffffll : (int64_t)v;
                else      x = (ovf || v > 0x8000000000000000ull) ? (int64_t)0x8000000000000000ull : (int64_t)(0ull - v);
                r.tlen = (int32_t)(uint32_t)(uint64_t)x;
            } else {                             // strtoul
                const uint64_t x = ovf ? ~0ull : (neg ? 0ull - v : v);
                if (conv == 1) r.flag = (uint32_t)x;
                if (conv == 3) r.pos = x;
                if (conv == 4) r.mapq = (uint32_t)x;
            }
        }
    }
    if (r.seq_len != qual_len) return kParseFail;    // sam-parse.c:50,88
    return kCounted;
}

// ---------------------------------------------------------------------------
// split_fast: a record [p0, pe) whose terminator sits at pe, located through
// `le`, the bit mask of bytes <= 0x20 (bit i of le[w] = byte 32w+i).  Accepts
// only clean records: ten '\t' between eleven non-empty fields, plain decimal
// numbers.  Everything else is handed to scan11 (return kNeedSlow).
//
// Written without early exits: on the GPU all lanes of a warp walk their
// records in lock step and re-converge after every loop; a lane that has seen
// something unusual just carries `ok = false` to the end.
// ---------------------------------------------------------------------------
// ---- decimal fields without byte loops -------------------------------------
// The digits of a field [a, e) are fetched as whole words ending at e (two or
// four aligned loads that do not depend on each other), bytes left of the
// field are replaced by '0', and the value comes out of a few multiply-adds.
// Compared with a load-test-multiply loop per digit this keeps the lanes of a
// warp in lock step and removes a chain of dependent shared-memory loads.
template <class B>
PSS_HD uint32_t word_at(const B &b, int p)            // bytes [p, p+4), any alignment, p >= 0
{
    const int a = p & ~3;
    return funnel_r(b.word(a), b.word(a + 4), 8u * (uint32_t)(p & 3));
}
PSS_HD bool all_digits4(uint32_t t) { return (((t + 0x76767676u) | t) & 0x80808080u) == 0u; }   // t = word ^ "0000"
PSS_HD uint32_t value4(uint32_t t)                    // four digit values, most significant in the low byte
{
    t = t * 10u + (t >> 8);
    return (t & 0xffu) * 100u + ((t >> 16) & 0xffu);
}
// w ^ "0000" with the first `drop` (>= 0) bytes forced to digit value 0: the digit values of a field that ends at the
// end of w.  One clamped funnel shift makes the byte mask (a shift by 32 or more gives 0), one LOP3 applies it.
PSS_HD uint32_t digit_tail(uint32_t w, int drop)
{
#if defined(__CUDA_ARCH__)
    const uint32_t m = __funnelshift_lc(0u, 0xffffffffu, 8u * (uint32_t)drop);
#else
    const uint32_t m = drop >= 4 ? 0u : (0xffffffffu << (8 * drop));
#endif
    return (w ^ 0x30303030u) & m;
}
template <class B>
PSS_HD uint32_t dec_loop(const B &b, int a, int e, bool &ok)      // only for the first bytes of the whole input
{
    uint32_t v = 0;
    for (int i = a; i < e; i++) { const uint32_t d = b(i) - '0'; ok = ok && (d <= 9u); v = v * 10u + d; }
    return v;
}
// value of the 1..4 digit field [a, e); anything else clears ok
template <class B>
PSS_HD uint32_t dec4(const B &b, int a, int e, bool &ok)
{
    int L = e - a;
    ok = ok && (L >= 1) && (L <= 4);
    if (!ok) L = 1;
    if (B::kLookBack < 4 && e - 4 < b.lo()) return dec_loop(b, e - L, e, ok);
    const uint32_t t = digit_tail(word_at(b, e - 4), 4 - L);
    ok = ok && all_digits4(t);
    return value4(t);
}
// [a, e) must be 1..9 decimal digits (value not needed); anything else clears ok
template <class B>
PSS_HD void digits9(const B &b, int a, int e, bool &ok)
{
    int L = e - a;
    ok = ok && (L >= 1) && (L <= 9);
    if (!ok) L = 1;
    if (B::kLookBack < 12 && e - 12 < b.lo()) { (void)dec_loop(b, e - L, e, ok); return; }
    const int      p = e - 12, a0 = p & ~3, drop = 12 - L;
    const uint32_t sh = 8u * (uint32_t)(p & 3);
    const uint32_t x0 = b.word(a0), x1 = b.word(a0 + 4), x2 = b.word(a0 + 8), x3 = b.word(a0 + 12);
    const uint32_t t0 = digit_tail(funnel_r(x0, x1, sh), drop);
    const uint32_t t1 = digit_tail(funnel_r(x1, x2, sh), drop - 4 < 0 ? 0 : drop - 4);
    const uint32_t t2 = digit_tail(funnel_r(x2, x3, sh), drop - 8 < 0 ? 0 : drop - 8);
    ok = ok && all_digits4(t0) && all_digits4(t1) && all_digits4(t2);
}
// value of the 1..9 digit field [a, e); anything else clears ok
template <class B>
PSS_HD uint32_t dec9(const B &b, int a, int e, bool &ok)
{
    int L = e - a;
    ok = ok && (L >= 1) && (L <= 9);
    if (!ok) L = 1;
    if (B::kLookBack < 12 && e - 12 < b.lo()) return dec_loop(b, e - L, e, ok);
    const int      p = e - 12, a0 = p & ~3, drop = 12 - L;          // drop in 3..11
    const uint32_t sh = 8u * (uint32_t)(p & 3);
    const uint32_t x0 = b.word(a0), x1 = b.word(a0 + 4), x2 = b.word(a0 + 8), x3 = b.word(a0 + 12);
    const uint32_t t0 = digit_tail(funnel_r(x0, x1, sh), drop);
    const uint32_t t1 = digit_tail(funnel_r(x1, x2, sh), drop - 4 < 0 ? 0 : drop - 4);
    const uint32_t t2 = digit_tail(funnel_r(x2, x3, sh), drop - 8 < 0 ? 0 : drop - 8);
    ok = ok && all_digits4(t0) && all_digits4(t1) && all_digits4(t2);
    return (value4(t0) * 10000u + value4(t1)) * 10000u + value4(t2);
}

template <class B>
PSS_HD int split_fast(const B &b, const uint32_t *le, int p0, int pe, RecView &r)
{
    int      sep[11];
    int      w = p0 >> 5;
    uint32_t bits = le[w] & (0xffffffffu << (p0 & 31));
    int      prev = p0 - 1;
    bool     ok = true;
    // next separator at or after the cursor (w, bits); empty mask words are skipped four at a time with
    // independent loads.  pe's bit is always set and 8 all-ones sentinel words follow the data.
#define PSS_NEXT_SEP(f)                                                                         \
    do {                                                                                        \
        while (bits == 0) {                                                                     \
            const uint32_t b1 = le[w + 1], b2 = le[w + 2], b3 = le[w + 3], b4 = le[w + 4];      \
            const int      k = b1 ? 1 : b2 ? 2 : b3 ? 3 : 4;                                    \
            bits = b1 ? b1 : b2 ? b2 : b3 ? b3 : b4;                                            \
            w += k;                                                                             \
        }                                                                                       \
        int p = (w << 5) + ffs32(bits) - 1;                                                     \
        bits &= bits - 1;                                                                       \
        if (p > pe) p = pe;                                                                     \
        ok = ok && (p > prev + 1);                 /* no empty field, no leading separator */   \
        if ((f) < 10) ok = ok && (p < pe);         /* fewer than 11 clean fields: scan11 decides */ \
        sep[f] = p;                                                                             \
        prev = p;                                                                               \
    } while (0)
    {   // end of QNAME: names of up to ~64 bytes end within three mask words, looked at in one go (longer: the loop)
        const uint32_t b1 = le[w + 1], b2 = le[w + 2];
        const bool     h0 = bits != 0u, h1 = b1 != 0u;
        const uint32_t x = h0 ? bits : h1 ? b1 : b2;
        w += h0 ? 0 : h1 ? 1 : 2;
        bits = x;
    }
    PSS_NEXT_SEP(0);                                // end of QNAME
    {
        // FLAG .. TLEN are short.  A 96-bit window of the mask, anchored at the byte after the last separator, is
        // shifted along: the next separator is then always the lowest set bit of its first word (fields of up to
        // 31 bytes; longer ones are left to scan11) -- one find-first-set and three funnel shifts per field, in
        // place of a chain of compares and selects over three words (the integer ALU pipe is what this kernel is
        // bound by).
        int            cur = sep[0] + 1;
        const int      w0 = cur >> 5;
        const uint32_t s0 = (uint32_t)(cur & 31);
        const uint32_t q0 = le[w0], q1 = le[w0 + 1], q2 = le[w0 + 2], q3 = le[w0 + 3];
        uint32_t h0 = funnel_r(q0, q1, s0), h1 = funnel_r(q1, q2, s0), h2 = funnel_r(q2, q3, s0);
#pragma unroll
        for (int f = 1; f <= 8; f++) {
            const int t = ffs32(h0);                // 1 + distance to the separator, 0: none within 32 bytes
            ok = ok && (t > 1);                     // found, and the field is not empty
            cur += t;
            sep[f] = cur - 1;
#if !defined(__CUDA_ARCH__)
            if (sep[f] > pe) sep[f] = pe;           // host build: stay inside the caller's buffer (the kernel's tile has slack; ok is false either way)
#endif
            h0 = funnel_rc(h0, h1, (uint32_t)t);
            h1 = funnel_rc(h1, h2, (uint32_t)t);
            h2 = funnel_rc(h2, 0u, (uint32_t)t);
        }
        ok = ok && (sep[8] < pe);                   // fewer than 11 clean fields: scan11 decides
        prev = sep[8];
        w = cur >> 5;                               // cursor for the two long fields
        bits = le[w] & (0xffffffffu << (cur & 31));
    }
    // SEQ and QUAL: reads of up to ~150 bases end within the next five mask words; looked at in one go, so that the
    // lanes of a warp do not loop a different number of times (longer fields fall through to the loop in PSS_NEXT_SEP)
#define PSS_SKIP5()                                                                             \
    do {                                                                                        \
        const uint32_t b1 = le[w + 1], b2 = le[w + 2], b3 = le[w + 3], b4 = le[w + 4], b5 = le[w + 5]; \
        const bool     h0 = bits != 0u, h1 = b1 != 0u, h2 = b2 != 0u, h3 = b3 != 0u, h4 = b4 != 0u;    \
        const uint32_t x = h0 ? bits : h1 ? b1 : h2 ? b2 : h3 ? b3 : h4 ? b4 : b5;              \
        w += h0 ? 0 : h1 ? 1 : h2 ? 2 : h3 ? 3 : h4 ? 4 : 5;                                    \
        bits = x;                                                                               \
    } while (0)
    PSS_SKIP5();
    PSS_NEXT_SEP(9);                                // end of SEQ
    PSS_SKIP5();
    PSS_NEXT_SEP(10);                               // end of QUAL
#undef PSS_SKIP5
#undef PSS_NEXT_SEP
    // separators 1..10 must be '\t', the 11th any white space (or the end of the line)
#pragma unroll
    for (int f = 0; f < 10; f++) ok = ok && (b(sep[f]) == '\t');
    ok = ok && (sep[10] == pe || is_ws(b(sep[10])));

    const int qname_len = sep[0] - p0;
    r.rname_off = sep[1] + 1; r.rname_len = sep[2] - sep[1] - 1;
    r.cigar_off = sep[4] + 1; r.cigar_len = sep[5] - sep[4] - 1;
    const int mrnm_len = sep[6] - sep[5] - 1;
    r.seq_off = sep[8] + 1;   r.seq_len = sep[9] - sep[8] - 1;
    const int qual_len = sep[10] - sep[9] - 1;
    ok = ok && ((qname_len | r.rname_len | r.cigar_len | mrnm_len | r.seq_len | qual_len) <= kMaxField);

    // plain decimal numbers of at most nine digits; anything else (signs on
    // unsigned fields, 0x.., leading-zero TLEN, longer numbers) goes to scan11
    r.flag = dec4(b, sep[0] + 1, sep[1], ok);                               // FLAG %u (more than 4 digits: scan11)
    r.pos = dec9(b, sep[2] + 1, sep[3], ok);                                // POS  %lu
    r.mapq = dec4(b, sep[3] + 1, sep[4], ok);                               // MAPQ %u
    digits9(b, sep[6] + 1, sep[7], ok);                                     // MPOS %u: never read, must convert cleanly
    {                                                                        // TLEN %i
        int a = sep[7] + 1;
        const int e = sep[8];
        const uint32_t c = b(a);
        const bool neg = (c == '-');
        if (c == '-' || c == '+') a++;
        ok = ok && !(e - a > 1 && b(a) == '0');    // a leading 0 would switch glibc to octal / hex
        const uint32_t v = dec4(b, a, e, ok);      // |TLEN| >= 10000 is left to scan11
        r.tlen = neg ? -(int32_t)v : (int32_t)v;
    }
    if (!ok) return kNeedSlow;
    if (r.seq_len != qual_len) return kParseFail;     // sam-parse.c:50,88
    return kCounted;
}

// ---------------------------------------------------------------------------
// genome access
// ---------------------------------------------------------------------------
// find_seq (fasta-genome-io.c:202-213): exact byte equality of RNAME and a
// contig id.  Returns contig index or -1.
template <class B>
PSS_HD int find_contig(const DevGenome &g, const B &b, int off, int len)
{
    uint32_t h = kNameHashSeed;
    for (int i = 0; i < len; i++) h = name_hash_step(h, (uint8_t)b(off + i));
    uint32_t slot = h & g.hash_mask;
    for (;;) {
        uint32_t v = g.hash[slot];
        if (v == 0) return -1;
        const DevContig &c = g.contigs[v - 1];
        if ((int)c.name_len == len) {
            const char *nm = g.names + c.name_off;
            int i = 0;
            while (i < len && (uint8_t)nm[i] == (uint8_t)b(off + i)) i++;
            if (i == len) return (int)(v - 1);
        }
        slot = (slot + 1) & g.hash_mask;
    }
}

// n_fields (<= 32) symbols starting at global base index gb: 2-bit codes and
// 2-bit classes, field j at bits 2j.  In two steps: fetch_window() issues the
// loads (HBM latency, a random 32-byte sector or two), window_fields() shifts
// the symbols into place.  `late` must be 0; the caller derives it from values
// it computes after the fetch (the decoded read bases), which keeps the
// compiler from scheduling the first use of the loaded words -- where a warp
// waits for HBM -- ahead of that work.
struct RawWindow {
    uint64_t g0, g1, g2;
    uint32_t sh;
};
PSS_HD RawWindow fetch_window(const DevGenome &g, uint64_t gb, int n_fields)
{
    RawWindow w;
    const uint64_t gi = gb >> 4;
    const uint32_t o = (uint32_t)(gb & 15u);
    w.sh = 2 * o;
    w.g0 = g.groups[gi];
    w.g1 = g.groups[gi + 1];
    w.g2 = (o + (uint32_t)n_fields > 32u) ? g.groups[gi + 2] : 0ull;
    return w;
}
PSS_HD void window_fields(const RawWindow &w, int n_fields, uint32_t late, uint64_t &codes, uint64_t &classes)
{
    const uint32_t sh = w.sh | late;
    const uint32_t c_lo = funnel_r((uint32_t)w.g0, (uint32_t)w.g1, sh);
    const uint32_t c_hi = funnel_r((uint32_t)w.g1, (uint32_t)w.g2, sh);
    const uint32_t k_lo = funnel_r((uint32_t)(w.g0 >> 32), (uint32_t)(w.g1 >> 32), sh);
    const uint32_t k_hi = funnel_r((uint32_t)(w.g1 >> 32), (uint32_t)(w.g2 >> 32), sh);
    const uint64_t m = low_fields_mask(n_fields);
    codes   = (((uint64_t)c_hi << 32) | c_lo) & m;
    classes = (((uint64_t)k_hi << 32) | k_lo) & m;
}
PSS_HD void load_window(const DevGenome &g, uint64_t gb, int n_fields, uint64_t &codes, uint64_t &classes)
{
    window_fields(fetch_window(g, gb, n_fields), n_fields, 0u, codes, classes);
}

// the byte behind an "other" symbol (binary search in the exception list)
PSS_HD uint8_t other_char(const DevGenome &g, uint64_t gb)
{
    uint32_t lo = 0, hi = g.n_exc;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        uint64_t p = g.exc_pos[mid];
        if (p == gb) return g.exc_chr[mid];
        if (p < gb) lo = mid + 1; else hi = mid;
    }
    return 0xff;
}

// strchr(UP_CTX|DOWN_CTX, c) != NULL for the genome symbol at gb, after the
// optional complement of a reverse-strand read (pss-bam.c:60-79 leaves
// non-ACGT bytes as they are).
PSS_HD bool ctx_member(const DevGenome &g, const TallyCfg &P, bool down, uint32_t sym, bool complement, uint64_t gb, bool live)
{
    if (sym < 4u && complement) sym ^= 3u;
    const uint32_t mask = down ? P.down_mask : P.up_mask;
    if (sym != kSymOther || !live) return ((mask >> (sym & 15u)) & 1u) != 0u;     // bit 15 is never set
    if (!(down ? P.down_other : P.up_other)) return false;
    const uint8_t c = other_char(g, gb);
    const char *s = down ? P.down_ctx : P.up_ctx;
    for (int i = 0; i < kMaxCtxChars && s[i]; i++)
        if ((uint8_t)s[i] == c) return true;
    return false;
}

// CIGAR must be the bytes of printf("%dM", n)  (pss-bam.c:113-123, fragkon.c:68-78)
template <class B>
PSS_HD bool cigar_is_nM(const B &b, int off, int len, int64_t n)
{
    bool ok = (len >= 2) && (len <= 10) && (n >= 0);
    if (!ok) len = 2;
    ok = ok && (b(off + len - 1) == 'M') && !(b(off) == '0' && len != 2);
    if (n >= 10000) {                     // only a paired record with a huge TLEN gets here: a real (rare) branch
        const uint32_t v = dec9(b, off, off + len - 1, ok);
        return ok && (int64_t)v == n;
    }
    // n has at most four digits and leading zeros are out: a longer CIGAR cannot be "<n>M".  (Records with indels have
    // longer CIGARs, and they are common: no second, nine-digit parse on their account.)
    ok = ok && (len <= 5);
    if (!ok) len = 2;
    const uint32_t v = dec4(b, off, off + len - 1, ok);
    return ok && (int64_t)v == n;
}

// ---------------------------------------------------------------------------
// read bases -> 2-bit codes, four at a time
// ---------------------------------------------------------------------------
PSS_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t out = 0;
    for (int i = 0; i < 4; i++) out |= (uint32_t)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return out;
#endif
}
// four ASCII bases (little endian in w) after toupper (pss-bam.c:425): 8 bits
// of codes (A=0 C=1 G=2 T=3, base k at bits 2k) and 8 bits with bit 2k set
// when base k is not one of ACGT (pss-bam.c:253-255 skips those cells).
PSS_HD void codes_of_word(uint32_t w, uint32_t &codes8, uint32_t &bad8)
{
    const uint32_t c = w | 0x20202020u;
    const uint32_t k = (c >> 1) & 0x03030303u;                 // a->0 c->1 g->3 t->2
    uint32_t sel = (k | (k >> 4)) & 0x00330033u;
    sel = (sel | (sel >> 8)) & 0x3333u;                        // one selector nibble per base
    const uint32_t x = c ^ byte_perm(0x67746361u, 0u, sel);    // "actg"[k]: zero byte <=> a valid base
    const uint32_t nz = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
    const uint32_t code = k ^ ((k >> 1) & 0x01010101u);        // a->0 c->1 g->2 t->3
    codes8 = (code * 0x01041040u) >> 24;
    bad8 = ((nz >> 7) * 0x01041040u) >> 24;
}
// 4*n_words (<= 32) read bases starting at byte `off`: field k = base k
template <class B, int NWC = 0>      // NWC: n_words as a compile-time constant (0: run time)
PSS_HD void read_codes(const B &b, int off, int n_words, uint64_t &codes, uint64_t &bad)
{
    if (NWC) n_words = NWC;
    codes = 0; bad = 0;
    const int      a0 = off & ~3;
    const uint32_t sh = 8u * (uint32_t)(off & 3);
    uint32_t       prev = b.word(a0);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k >= n_words) break;
        const uint32_t next = b.word(a0 + 4 * k + 4);
        uint32_t c8, b8;
        codes_of_word(funnel_r(prev, next, sh), c8, b8);
        codes |= (uint64_t)c8 << (8 * k);
        bad   |= (uint64_t)b8 << (8 * k);
        prev = next;
    }
}

// What one record contributes to the two count tables, in table-row order:
// field j of *_ref/*_read = row j (rows 0,1 = the context bases two / one
// away, read := ref there; row i+2 = position i from that end).  A set bit 2j
// of *_bad means "row j adds nothing".
struct PssStreams {
    uint64_t a_ref, a_read, a_bad;     // -> fwd_counts (5' end of the molecule)
    uint64_t b_ref, b_read, b_bad;     // -> rev_counts (3' end)
};

constexpr uint64_t kEvenBits = 0x5555555555555555ull;

// pss-bam.c:390-496 process_aln, after find_seq (:393) has been resolved by
// the caller: ci < 0 means "no such contig".  Fills `st` (all-bad unless
// counted).
//
// Written as straight-line predicated code: a record that fails a filter
// keeps walking with `live == false` and harmless addresses instead of
// returning, so that the 32 lanes of a warp stay converged through the window
// gathers and the read decoding (a lane that drops out costs nothing, a warp
// that splits executes everything twice).  `valid` = the record parsed; lanes
// without a record pass valid == false and the RecView defaults.
template <class B, int RC = -1>      // RC: -r as a compile-time constant (-1: taken from P)
PSS_HD int pss_record(const B &b, const RecView &r, bool valid, int ci, uint64_t ctg_base, uint64_t ctg_len,
                      const DevGenome &g, const TallyCfg &P, PssStreams &st)
{
    int  code = kCounted;
    bool live = valid;
#define PSS_DROP(cond, c) do { if (live && (cond)) { code = (c); live = false; } } while (0)
    PSS_DROP(ci < 0, kNoContig);                                            // :393-396
    PSS_DROP(ctg_len == 0, kUndefined);           // `ref->len-1` wraps (:408) and the window copy reads past the string

    const bool paired = r.flag & 1u;
    // sam-parse.c:66-68: unpaired -> isize = strlen(seq); pss-bam.c:401: n = abs(isize)
    const uint32_t abs_tlen = r.tlen < 0 ? 0u - (uint32_t)r.tlen : (uint32_t)r.tlen;       // |INT_MIN| fits
    const int64_t  n = (int64_t)(paired ? abs_tlen : (uint32_t)r.seq_len);
    PSS_DROP(n > kMaxTlen, kUndefined);           // reference: stack overflow in its VLAs before any filter
    const int     R = RC >= 0 ? RC : P.R;
    const int64_t s = (int64_t)(r.pos - 1);       // :403
    const int64_t e = s + n - 1;                  // :404

    PSS_DROP(s - 2 < 0, kFiltered);                                         // :407
    PSS_DROP((uint64_t)(e + 2) > ctg_len - 1, kFiltered);                   // :408
    PSS_DROP(r.mapq < P.min_mq, kFiltered);                                 // :409
    PSS_DROP(!((uint64_t)n >= P.min_len && (uint64_t)n <= P.max_len && n >= R), kFiltered);   // :96-103
    PSS_DROP(r.flag & (4u | 256u | 512u | 1024u | 2048u), kFiltered);       // :412-416
    PSS_DROP(P.merged_only && paired, kFiltered);                           // :417

    // the two windows leave for HBM before the rest of the record is looked at
    const int W = R + 2;
    uint64_t lc, lk, rc, rk;
    const RawWindow lw = fetch_window(g, live ? ctg_base + (uint64_t)(s - 2) : (uint64_t)kPadBases, W);             // g[0 .. W)
    const RawWindow rw = fetch_window(g, live ? ctg_base + (uint64_t)(e + 2 - (W - 1)) : (uint64_t)kPadBases, W);   // g[n+3-(W-1) .. n+3]

    PSS_DROP(!cigar_is_nM(b, r.cigar_off, r.cigar_len, n), kFiltered);      // :411
    // paired: n comes from TLEN; with a shorter SEQ the reference reads bytes
    // of earlier records that are still in its Saml buffer
    PSS_DROP(paired && (int64_t)r.seq_len < n, kUndefined);

    // read bases: r[0..R) and, reversed, r[n-1-i]; both moved up two rows
    const int nw = (R + 3) >> 2;
    const int nn = live ? (int)n : 4 * nw;
    uint64_t  pre, pre_bad, suf, suf_bad;
    read_codes<B, (RC >= 0 ? (RC + 3) >> 2 : 0)>(b, r.seq_off, nw, pre, pre_bad);
    read_codes<B, (RC >= 0 ? (RC + 3) >> 2 : 0)>(b, r.seq_off + nn - 4 * nw, nw, suf, suf_bad);
    suf = rev_fields64(suf, 4 * nw);
    suf_bad = rev_fields64(suf_bad, 4 * nw);
    const uint64_t m = low_fields_mask(W);
    pre = (pre << 4) & m;  pre_bad = (pre_bad << 4) & m;
    suf = (suf << 4) & m;  suf_bad = (suf_bad << 4) & m;

    // the genome windows are first touched here, after the read has been decoded (P.zero is 0)
    const uint32_t late = ((uint32_t)pre ^ (uint32_t)suf ^ (uint32_t)pre_bad ^ (uint32_t)suf_bad) & P.zero;
    window_fields(lw, W, late, lc, lk);
    window_fields(rw, W, late, rc, rk);

    rc = rev_fields64(rc, W);                                               // row j <-> g[n+3-j]
    rk = rev_fields64(rk, W);

    // symbols of the two adjacent context bases (row 1 of each side)
    const bool     rev = r.flag & 16u;
    const uint64_t gb_up = ctg_base + (uint64_t)(s - 1);   // g[1]
    const uint64_t gb_dn = ctg_base + (uint64_t)(e + 1);   // g[n+2]
    const uint32_t sym_l = (uint32_t)((lc >> 2) & 3u) | ((uint32_t)((lk >> 2) & 3u) << 2);
    const uint32_t sym_r = (uint32_t)((rc >> 2) & 3u) | ((uint32_t)((rk >> 2) & 3u) << 2);
    // molecule orientation: forward read -> 5' context is g[1]; reverse read ->
    // 5' context is comp(g[n+2]) (pss-bam.c:430-436)
    const bool up_ok = ctx_member(g, P, false, rev ? sym_r : sym_l, rev, rev ? gb_dn : gb_up, live);
    const bool dn_ok = ctx_member(g, P, true, rev ? sym_l : sym_r, rev, rev ? gb_up : gb_dn, live);

    // which table(s) this record feeds: unpaired :428-447, paired :450-493
    const bool pp = ((r.flag & 2u) != 0u) & ((r.flag & 8u) == 0u);
    const bool sel_a = pp & ((r.flag & 64u) != 0u) & up_ok;                 // :460 / :482
    const bool sel_b = pp & !sel_a & ((r.flag & 128u) != 0u) & dn_ok;       // :471 / :488
    const bool want_a = paired ? sel_a : (up_ok & dn_ok);
    const bool want_b = paired ? sel_b : (up_ok & dn_ok);
    PSS_DROP(!(want_a | want_b), kFiltered);
#undef PSS_DROP

    // class != 0 -> not one of ACGT -> the cell is skipped (:253-255, :176-188)
    const uint64_t l_bad = ((lk | (lk >> 1)) & kEvenBits) | pre_bad;
    const uint64_t r_bad = ((rk | (rk >> 1)) & kEvenBits) | suf_bad;
    const uint64_t l_read = (lc & 0xfull) | pre;      // rows 0,1: diagonal cell of the context base
    const uint64_t r_read = (rc & 0xfull) | suf;
    const uint64_t flip = rev ? m : 0ull;             // complement = flip both code bits
    st.a_ref  = (rev ? rc : lc) ^ flip;
    st.a_read = (rev ? r_read : l_read) ^ flip;
    st.a_bad  = (live && want_a) ? (rev ? r_bad : l_bad) : kEvenBits;
    st.b_ref  = (rev ? lc : rc) ^ flip;
    st.b_read = (rev ? l_read : r_read) ^ flip;
    st.b_bad  = (live && want_b) ? (rev ? l_bad : r_bad) : kEvenBits;
    return code;
}

// ---------------------------------------------------------------------------
// pss-bam.c:390-496 for any -r (R > kMaxRegion does not fit the 64-bit row
// streams above): the same filters in the same order, then one table cell per
// row and side handed to add(table, row, cell) -- table 0 = fwd_counts (5'
// end), 1 = rev_counts.  Base by base, with real branches: -r beyond 30 is an
// unusual request and this path only has to be exact.
// ---------------------------------------------------------------------------
PSS_HD uint32_t genome_sym(const DevGenome &g, uint64_t gb)
{
    const uint64_t grp = g.groups[gb >> 4];
    const uint32_t sh = 2u * (uint32_t)(gb & 15u);
    return (((uint32_t)grp >> sh) & 3u) | ((((uint32_t)(grp >> 32) >> sh) & 3u) << 2);
}
// code 0..3 of a read base after toupper (pss-bam.c:425), 4 if it is not one of ACGT
PSS_HD uint32_t read_base_code(uint32_t c)
{
    c |= 0x20u;
    return c == 'a' ? 0u : c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 4u;
}
template <class B, class Add>
PSS_HD int pss_record_wide(const B &b, const RecView &r, int ci, uint64_t ctg_base, uint64_t ctg_len,
                           const DevGenome &g, const TallyCfg &P, Add add)
{
    if (ci < 0) return kNoContig;                                           // :393-396
    if (ctg_len == 0) return kUndefined;
    const bool    paired = r.flag & 1u;
    const uint32_t abs_tlen = r.tlen < 0 ? 0u - (uint32_t)r.tlen : (uint32_t)r.tlen;       // |INT_MIN| fits
    const int64_t  n = (int64_t)(paired ? abs_tlen : (uint32_t)r.seq_len);
    if (n > kMaxTlen) return kUndefined;
    const int     R = P.R;
    const int64_t s = (int64_t)(r.pos - 1), e = s + n - 1;                  // :403-404
    if (s - 2 < 0) return kFiltered;                                        // :407
    if ((uint64_t)(e + 2) > ctg_len - 1) return kFiltered;                  // :408
    if (r.mapq < P.min_mq) return kFiltered;                                // :409
    if (!((uint64_t)n >= P.min_len && (uint64_t)n <= P.max_len && n >= R)) return kFiltered;   // :96-103
    if (r.flag & (4u | 256u | 512u | 1024u | 2048u)) return kFiltered;      // :412-416
    if (P.merged_only && paired) return kFiltered;                          // :417
    if (!cigar_is_nM(b, r.cigar_off, r.cigar_len, n)) return kFiltered;     // :411
    if (paired && (int64_t)r.seq_len < n) return kUndefined;

    const bool     rev = r.flag & 16u;
    const uint64_t g0 = ctg_base + (uint64_t)(s - 2);                       // g[0]
    const uint64_t gN = ctg_base + (uint64_t)(e + 2);                       // g[n+3]
    const bool up_ok = ctx_member(g, P, false, genome_sym(g, rev ? gN - 1 : g0 + 1), rev, rev ? gN - 1 : g0 + 1, true);
    const bool dn_ok = ctx_member(g, P, true, genome_sym(g, rev ? g0 + 1 : gN - 1), rev, rev ? g0 + 1 : gN - 1, true);
    const bool pp = (r.flag & 2u) && !(r.flag & 8u);
    const bool sel_a = pp && (r.flag & 64u) && up_ok;                       // :460 / :482
    const bool sel_b = pp && !sel_a && (r.flag & 128u) && dn_ok;            // :471 / :488
    const bool want_a = paired ? sel_a : (up_ok && dn_ok);
    const bool want_b = paired ? sel_b : (up_ok && dn_ok);
    if (!(want_a || want_b)) return kFiltered;

    const uint32_t flip = rev ? 3u : 0u;                                    // complement
    for (int j = 0; j < R + 2; j++) {
        // row j: context rows 0, 1 (read := ref), then position j - 2 from that end
        const uint32_t sl = genome_sym(g, g0 + (uint64_t)j), sr = genome_sym(g, gN - (uint64_t)j);
        const uint32_t ql = j < 2 ? sl : read_base_code(b(r.seq_off + (j - 2)));
        const uint32_t qr = j < 2 ? sr : read_base_code(b(r.seq_off + (int)n - 1 - (j - 2)));
        const uint32_t a_ref = rev ? sr : sl, a_read = rev ? qr : ql;       // 5' end of the molecule
        const uint32_t b_ref = rev ? sl : sr, b_read = rev ? ql : qr;       // 3' end
        if (want_a && a_ref < 4u && a_read < 4u) add(0, j, (int)(((a_read ^ flip) << 2) | (a_ref ^ flip)));
        if (want_b && b_ref < 4u && b_read < 4u) add(1, j, (int)(((b_read ^ flip) << 2) | (b_ref ^ flip)));
    }
    return kCounted;
}

// ---------------------------------------------------------------------------
// fragkon.c:122-216 process_aln.  Outputs up to two histogram indices
// (MSB-first 2-bit k-mer code, kmer.c:184-214); *_ok says whether to count.
// ---------------------------------------------------------------------------
struct FkHits {
    uint32_t idx5, idx3;
    bool     add5, add3;
};

// K (<= 14) symbols starting at gb -> LSB-first code word; false if any is not ACGT
PSS_HD bool load_kmer(const DevGenome &g, uint64_t gb, int K, uint32_t &lsb_first)
{
    const uint64_t gi = gb >> 4;
    const uint32_t sh = 2 * (uint32_t)(gb & 15u);
    const uint64_t g0 = g.groups[gi], g1 = g.groups[gi + 1];
    const uint32_t m = (K >= 16) ? 0xffffffffu : ((1u << (2 * K)) - 1u);
    lsb_first = funnel_r((uint32_t)g0, (uint32_t)g1, sh) & m;
    return (funnel_r((uint32_t)(g0 >> 32), (uint32_t)(g1 >> 32), sh) & m) == 0u;
}
PSS_HD uint32_t kmer_index_fwd(uint32_t lsb_first, int K) { return rev_fields32(lsb_first) >> (32 - 2 * K); }
// reverse complement: the LSB-first word of the forward strand IS the
// MSB-first word of the reversed k-mer; complement = flip all bits
PSS_HD uint32_t kmer_index_rc(uint32_t lsb_first, int K) { return ~lsb_first & ((1u << (2 * K)) - 1u); }

template <class B>
PSS_HD int fk_record(const B &b, const RecView &r, bool valid, int ci, uint64_t ctg_base, uint64_t ctg_len,
                     const DevGenome &g, const TallyCfg &P, FkHits &h)
{
    int  code = kFiltered;
    bool live = valid;
#define PSS_DROP(cond, c) do { if (live && (cond)) { code = (c); live = false; } } while (0)
    PSS_DROP(ci < 0, kNoContig);                                           // :124-127
    PSS_DROP(ctg_len == 0, kUndefined);           // `ref->len-1` wraps (:138)

    const int      K = P.K;
    const uint64_t ok = (uint64_t)(K / 2), ik = (uint64_t)K - ok;          // :134-135
    const uint64_t n = (uint64_t)r.seq_len;                                 // :130 (strlen(seq), not TLEN)
    const uint64_t s = r.pos - 1;                                           // :129, unsigned: POS 0 wraps
    const uint64_t e = s + n - 1;
    // `aln_start-(KLEN/2) >= 0` is an unsigned tautology (:137)
    PSS_DROP(!(e + ok <= ctg_len - 1), kFiltered);                         // :138
    PSS_DROP(!(r.mapq >= P.min_mq), kFiltered);                            // :139
    PSS_DROP(!(n >= P.min_len && n <= P.max_len), kFiltered);              // :52-58
    PSS_DROP(r.flag & (4u | 256u | 512u | 1024u | 2048u), kFiltered);      // :142-146

    // Window starts.  Forward: 5' G[s-ok, +K), 3' G[s+n-ik, +K) (:176-177).
    // Reverse: sub = G[s-ok, s-ok+n+K) copied with strncpy (:156), 5' =
    // RC(sub)[0,K) = RC(G[s-ok+n, +K)) (:164), 3' = RC(sub)[ok+n-ik, +K) =
    // RC(G[s-ok+(ik-ok), +K)) (:167).  Left of the contig the reference reads
    // the allocator's header, whose last byte is 0: such a window never
    // validates (the packed genome has "other" symbols there), and for reverse
    // reads the strncpy stops at that 0 and zero-fills, so s < ok kills both.
    const bool    rev = r.flag & 16u;
    const int64_t ss = (int64_t)s;                // -1 when POS was 0
    const int64_t o5 = rev ? ss - (int64_t)ok + (int64_t)n : ss - (int64_t)ok;
    const int64_t o3 = rev ? ss - (int64_t)ok + (int64_t)(ik - ok) : ss + (int64_t)n - (int64_t)ik;
    uint32_t w5, w3;
    bool v5 = load_kmer(g, live ? ctg_base + (uint64_t)o5 : (uint64_t)kPadBases, K, w5);
    bool v3 = load_kmer(g, live ? ctg_base + (uint64_t)o3 : (uint64_t)kPadBases, K, w3);
    if (rev && ss < (int64_t)ok) v5 = v3 = false;
    h.idx5 = rev ? kmer_index_rc(w5, K) : kmer_index_fwd(w5, K);
    h.idx3 = rev ? kmer_index_rc(w3, K) : kmer_index_fwd(w3, K);

    PSS_DROP(!cigar_is_nM(b, r.cigar_off, r.cigar_len, (int64_t)n), kFiltered);   // :141
#undef PSS_DROP

    const bool paired = r.flag & 1u;
    const bool pp = !P.merged_only && (r.flag & 2u) && !(r.flag & 8u);     // :187-190
    const bool use5 = paired ? (pp && (r.flag & 64u)) : true;              // unpaired: both ends, independently (:149-184)
    const bool use3 = paired ? (pp && !(r.flag & 64u) && (r.flag & 128u)) : true;
    h.add5 = live && use5 && v5;
    h.add3 = live && use3 && v3;
    if (live) {
        if (!paired) code = (v5 && v3) ? kCounted : kFiltered;
        else         code = ((use5 && v5) || (use3 && v3)) ? kCounted : kFiltered;
    }
    return code;
}

}  // namespace pssgpu
