// pss_inflate.h -- raw DEFLATE (RFC 1951) decoder for BGZF blocks, written once for the sm_100a inflate kernel
// (pss_bam.cuh) and for a host build that the CPU test-suite checks against zlib (tests/host_emul).  The host
// build is a TEST of this logic; the product library only ever executes it inside a CUDA kernel.
//
// Replaces the `samtools view` child of the reference (pss-bam.c:148-162, fragkon.c:84-93), whose first stage is the
// BGZF inflate.  The BGZF container (SAM spec 4.1): a series of gzip members of at most 64 KiB of payload each,
// independent of one another -- which is what makes the format GPU friendly: one WARP per BGZF block.
//
// Execution model.  DEFLATE decoding is a serial chain (the position of a symbol is known only when the previous one
// has been decoded), so the 32 lanes of a warp run the symbol loop *redundantly* with identical, warp-uniform state:
// no divergence, no shuffles in the steady state, every table look-up is a shared-memory broadcast.  The lanes part
// ways only where there is parallel work:
//   * the compressed bytes are fetched 128 bytes at a time, one 32-bit word per lane (coalesced), and handed to the
//     bit buffer with one shuffle per 32 bits;
//   * LZ77 matches are copied by all lanes together (byte i by lane i mod 32);
//   * the decode tables of a dynamic block are filled by all lanes together.
// Parallelism across the GPU comes from the ~10^5 independent BGZF blocks of a batch.
//
// Tables (per warp, in shared memory): 32-bit entries that carry everything the symbol loop needs -- the literal byte
// or the base of a length / distance together with the code length and "code length + extra bits" -- so that a symbol
// costs one look-up and a match needs no second table (RFC 1951 3.2.5 is folded in when a block's tables are built).
// Literal/length codes longer than the first level go through second-level tables behind it (zlib's scheme); the
// canonical count/sorted-symbol arrays, walked bit by bit, remain as the fallback (second-level area full, long distance
// codes).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PSS_IHD __host__ __device__ __forceinline__
#define PSS_IHD_COLD __host__ __device__ __noinline__      // set-up and rare paths: kept out of the symbol loop's register budget
#else
#define PSS_IHD inline
#define PSS_IHD_COLD inline
#endif

namespace pssgpu {

#ifndef PSS_INF_LIT_BITS
#define PSS_INF_LIT_BITS 10
#endif
constexpr int kInfLitBits  = PSS_INF_LIT_BITS;               // first-level bits of the literal/length table
constexpr int kInfDistBits = 8;
#ifndef PSS_INF_LIT_SUB
#define PSS_INF_LIT_SUB (PSS_INF_LIT_BITS >= 10 ? 288 : 352)
#endif
constexpr int kInfLitSub   = PSS_INF_LIT_SUB;                // second-level entries (zlib's worst case needs 308 / 340; a block that
                                                             // needs more than there are takes the canonical walk for the rest)
constexpr int kInfMaxLit   = 288;
constexpr int kInfMaxDist  = 32;

// table entry: bits 0..4 code length (second level: of the whole code), bit 5 literal, bit 6 second-level pointer,
// bit 7 special, bits 8..12 code length + extra bits (pointer: bits of the second-level index), bits 16..31 value:
// literal byte / base length / base distance / start of the second-level table / kind of special
constexpr uint32_t kInfELit = 0x20u, kInfESub = 0x40u, kInfESpecial = 0x80u;
constexpr uint32_t kInfSpEnd = 0u, kInfSpWalk = 1u, kInfSpInvalid = 2u;      // special: end of block / canonical walk / no such code

struct InflateTables {                       // one per warp
    uint32_t lit_lut[(1 << kInfLitBits) + kInfLitSub];
    uint32_t dist_lut[1 << kInfDistBits];
    uint16_t lit_sorted[kInfMaxLit];         // symbols in canonical order (by code length, then symbol)
    uint16_t dist_sorted[kInfMaxDist];
    uint16_t lit_count[16], dist_count[16];  // codes per length
    uint32_t lit_first, lit_index;           // state of the canonical walk after kInfLitBits levels (longer codes resume there)
    uint32_t dist_first, dist_index;         // ... after kInfDistBits levels
    uint8_t  lens[kInfMaxLit + kInfMaxDist]; // code lengths of the block being set up
    uint32_t window[32];                     // the current 32 words of compressed input (InfBits)
    uint32_t stage[32];                      // per lane: the aligned word around the source byte of the match in flight (inf_copy)
};

enum : int {
    kInfOk = 0,
    kInfBadBlockType = 1, kInfBadStored = 2, kInfBadCodeLengths = 3, kInfBadSymbol = 4, kInfBadDistance = 5,
    kInfOutputOverrun = 6, kInfInputOverrun = 7, kInfSizeMismatch = 8,
    kInfCrcMismatch = 9                                      // set by the inflate kernel (pss_crc32.h), not by inflate_block
};

// ---- lane model --------------------------------------------------------------------------------------------------
// Device: 32 lanes.  Host: one lane that plays all of them (the cooperative loops degenerate to plain loops).
struct InfLanes {
#if defined(__CUDA_ARCH__)
    static __device__ __forceinline__ int  lane() { return (int)(threadIdx.x & 31u); }
    static __device__ __forceinline__ int  width() { return 32; }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    static __device__ __forceinline__ uint32_t bcast(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#else
    static int  lane() { return 0; }
    static int  width() { return 1; }
    static void sync() {}
    static uint32_t bcast(uint32_t v, int) { return v; }
#endif
};

// ---- bit reader -----------------------------------------------------------------------------------------------------
// The compressed bytes are read as aligned 32-bit words starting at the word that holds the first byte.  On the device
// the current 32-word window sits in the warp's shared memory (a refill is one broadcast load, no shuffle), lane l
// holds word l of the next window in a register -- already in flight -- and stores it when the window is used up;
// `in_words` bounds the reads (words beyond it read as 0 and raise kInfInputOverrun only if their bits are consumed).
struct InfBits {
    const uint32_t *words;       // 4-byte aligned
    uint32_t        n_words;     // words that may be read
    uint32_t        next;        // index of the next word to enter the bit buffer
    uint64_t        buf;
    int             cnt;         // valid bits in buf
    uint32_t        win_next;    // device: this lane's word of the next 32-word window
    uint32_t        wbase;       // device: shared-space address of the current window

    PSS_IHD uint32_t load(uint32_t i) const
    {
#if defined(__CUDA_ARCH__)
        return i < n_words ? __ldg(words + i) : 0u;
#else
        return i < n_words ? words[i] : 0u;
#endif
    }
    PSS_IHD void open(const uint8_t *p, uint32_t n_bytes, uint32_t *window)
    {
        const uintptr_t a = (uintptr_t)p;
        const uint32_t  mis = (uint32_t)(a & 3u);
        words = reinterpret_cast<const uint32_t *>(a - mis);
        n_words = (n_bytes + mis + 3u) >> 2;
        next = 0;
#if defined(__CUDA_ARCH__)
        wbase = (uint32_t)__cvta_generic_to_shared(window);
        InfLanes::sync();                                    // nobody reads the old window any more
        window[InfLanes::lane()] = load((uint32_t)InfLanes::lane());
        win_next = load(32u + (uint32_t)InfLanes::lane());
        InfLanes::sync();
#else
        (void)window;
        win_next = wbase = 0;
#endif
        buf = 0; cnt = 0;
        const uint32_t w = take_word();
        buf = (uint64_t)(w >> (8u * mis));
        cnt = 32 - 8 * (int)mis;
    }
    PSS_IHD uint32_t take_word()
    {
#if defined(__CUDA_ARCH__)
        uint32_t w;
        asm volatile("{\n\t.reg .u32 t;\n\tmad.lo.u32 t, %1, 4, %2;\n\tld.shared.u32 %0, [t];\n\t}" : "=r"(w) : "r"(next & 31u), "r"(wbase) : "memory");
        next++;
        if ((next & 31u) == 0u) {                  // warp uniform: the window is used up, the next one moves in
            InfLanes::sync();
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(wbase + 4u * (uint32_t)InfLanes::lane()), "r"(win_next) : "memory");
            win_next = load(next + 32u + (uint32_t)InfLanes::lane());
            InfLanes::sync();
        }
        return w;
#else
        return load(next++);
#endif
    }
    // at least 33 valid bits afterwards
    PSS_IHD void refill()
    {
        if (cnt <= 32) {
            buf |= (uint64_t)take_word() << cnt;
            cnt += 32;
        }
    }
    PSS_IHD uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }      // n <= 16
    PSS_IHD void     drop(int n) { buf >>= n; cnt -= n; }
    PSS_IHD uint32_t get(int n) { const uint32_t v = peek(n); drop(n); return v; }
    // bytes consumed so far, counting a partly used byte as consumed
    PSS_IHD uint32_t bytes_used(uint32_t mis) const { return next * 4u - mis - (uint32_t)(cnt >> 3); }
};

// base | extra bits << 16 of the length symbols 257..285 and the distance symbols 0..29 (RFC 1951 3.2.5)
#if defined(__CUDACC__)
#define PSS_INF_CONST static __device__ __constant__
#else
#define PSS_INF_CONST static const
#endif
PSS_INF_CONST uint32_t kInfLenTab[32] = {
    3, 4, 5, 6, 7, 8, 9, 10, 11 | (1 << 16), 13 | (1 << 16), 15 | (1 << 16), 17 | (1 << 16), 19 | (2 << 16), 23 | (2 << 16), 27 | (2 << 16),
    31 | (2 << 16), 35 | (3 << 16), 43 | (3 << 16), 51 | (3 << 16), 59 | (3 << 16), 67 | (4 << 16), 83 | (4 << 16), 99 | (4 << 16),
    115 | (4 << 16), 131 | (5 << 16), 163 | (5 << 16), 195 | (5 << 16), 227 | (5 << 16), 258, 0, 0, 0 };
PSS_INF_CONST uint32_t kInfDistTab[32] = {
    1, 2, 3, 4, 5 | (1 << 16), 7 | (1 << 16), 9 | (2 << 16), 13 | (2 << 16), 17 | (3 << 16), 25 | (3 << 16), 33 | (4 << 16), 49 | (4 << 16),
    65 | (5 << 16), 97 | (5 << 16), 129 | (6 << 16), 193 | (6 << 16), 257 | (7 << 16), 385 | (7 << 16), 513 | (8 << 16), 769 | (8 << 16),
    1025 | (9 << 16), 1537 | (9 << 16), 2049 | (10 << 16), 3073 | (10 << 16), 4097 | (11 << 16), 6145 | (11 << 16), 8193 | (12 << 16),
    12289 | (12 << 16), 16385 | (13 << 16), 24577 | (13 << 16), 0, 0 };
#if defined(__CUDACC__) && !defined(__CUDA_ARCH__)
// host pass of nvcc: the device tables above cannot be read here; the host never decodes in the product library
PSS_IHD uint32_t inf_len_tab(int) { return 0; }
PSS_IHD uint32_t inf_dist_tab(int) { return 0; }
#else
PSS_IHD uint32_t inf_len_tab(int s) { return kInfLenTab[s]; }
PSS_IHD uint32_t inf_dist_tab(int s) { return kInfDistTab[s]; }
#endif

// ---- table entries ----------------------------------------------------------------------------------------------------
enum : int { kInfKindPlain = 0, kInfKindLit = 1, kInfKindDist = 2 };      // what the symbols of a code are
// the entry of symbol s with code length l
PSS_IHD uint32_t inf_entry(int kind, int s, int l)
{
    const uint32_t ul = (uint32_t)l;
    if (kind == kInfKindPlain) return ((uint32_t)s << 16) | ul;
    if (kind == kInfKindLit) {
        if (s < 256) return ((uint32_t)s << 16) | kInfELit | ul;
        if (s == 256) return (kInfSpEnd << 16) | kInfESpecial | ul;
        if (s > 285) return (kInfSpInvalid << 16) | kInfESpecial | ul;
        const uint32_t lc = inf_len_tab(s - 257);
        return ((lc & 0xffffu) << 16) | ((ul + (lc >> 16)) << 8) | ul;
    }
    if (s > 29) return (kInfSpInvalid << 16) | kInfESpecial | ul;
    const uint32_t dc = inf_dist_tab(s);
    return ((dc & 0xffffu) << 16) | ((ul + (dc >> 16)) << 8) | ul;
}
PSS_IHD uint32_t inf_brev(uint32_t c, int l)      // the l-bit code c in the order its bits arrive
{
#if defined(__CUDA_ARCH__)
    return __brev(c) >> (32 - l);
#else
    uint32_t r = 0;
    for (int b = 0; b < l; b++) r |= ((c >> b) & 1u) << (l - 1 - b);
    return r;
#endif
}

// ---- table set-up ---------------------------------------------------------------------------------------------------
// Canonical Huffman code of `n` symbols with code lengths lens[0..n): fills count[], sorted[] and the table lut:
// 2^bits first-level entries and, when sub_cap > 0, second-level tables for the longer codes in the sub_cap entries
// behind them.  First-level slots that no short code claims are left as "walk" entries (canonical walk: it finds the
// long codes no second-level table was built for, and rejects bit patterns that are no code at all).  Returns false
// for an over-subscribed or (non-trivially) incomplete code.  All lanes call it together; the symbol loops are uniform,
// the table fills are spread over the lanes.
PSS_IHD_COLD bool inf_build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sorted, uint32_t *lut, int bits, int kind, int sub_cap,
                            uint32_t *resume_first = nullptr, uint32_t *resume_index = nullptr)
{
    const int lane = InfLanes::lane(), W = InfLanes::width();
    for (int i = lane; i < 16; i += W) count[i] = 0;
    for (int i = lane; i < (1 << bits) + sub_cap; i += W) lut[i] = (kInfSpWalk << 16) | kInfESpecial;
    InfLanes::sync();
    if (lane == 0)
        for (int s = 0; s < n; s++) count[lens[s]]++;
    InfLanes::sync();
    // offsets / first codes per length (uniform)
    uint32_t offs[16], code[16];
    int      left = 1, max_len = 0;
    bool     ok = true;
    {
        uint32_t o = 0, c = 0;
        offs[0] = 0; code[0] = 0;
        for (int l = 1; l < 16; l++) {
            left <<= 1;
            left -= (int)count[l];
            if (left < 0) ok = false;
            offs[l] = o; code[l] = c;
            o += count[l];
            c = (c + count[l]) << 1;
            if (count[l]) max_len = l;
        }
    }
    if (resume_first && lane == 0) {           // canonical walk (inf_long) after `bits` levels without a hit
        uint32_t f = 0, ix = 0;
        for (int l = 1; l <= bits; l++) { ix += count[l]; f = (f + count[l]) << 1; }
        *resume_first = f;
        *resume_index = ix;
    }
    const int used = n - (int)count[0];
    // incomplete codes are legal only in the one-code case (a single distance code, RFC 1951 3.2.7) -- and zlib also
    // lets a block with no distance code at all pass
    if (left > 0 && used > 1) ok = false;
    if (!ok) return false;
    // second-level tables, first pass (uniform, no table reads): codes longer than `bits` in canonical order -- their
    // left-aligned values grow, so the codes that share their first `bits` bits are neighbours and the last of them is
    // the longest: one table of 2^(that length - bits) entries per distinct prefix
    if (sub_cap > 0 && max_len > bits) {
        int      cur_p = -1, cur_len = 0, next_free = 0;
        bool     full = false;
        for (int l = bits + 1; l <= max_len; l++) {
            for (uint32_t k = 0; k < count[l]; k++) {
                const int p = (int)(inf_brev(code[l] + k, l) & ((1u << bits) - 1u));
                if (p != cur_p) {
                    if (cur_p >= 0 && !full) {
                        const int sb = cur_len - bits;
                        if (next_free + (1 << sb) <= sub_cap) {
                            if (lane == 0) lut[cur_p] = ((uint32_t)((1 << bits) + next_free) << 16) | ((uint32_t)sb << 8) | kInfESub | (uint32_t)bits;
                            next_free += 1 << sb;
                        } else full = true;                    // this and the remaining prefixes stay "walk"
                    }
                    cur_p = p;
                }
                cur_len = l;
            }
        }
        if (cur_p >= 0 && !full) {
            const int sb = cur_len - bits;
            if (next_free + (1 << sb) <= sub_cap && lane == 0)
                lut[cur_p] = ((uint32_t)((1 << bits) + next_free) << 16) | ((uint32_t)sb << 8) | kInfESub | (uint32_t)bits;
        }
        InfLanes::sync();
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (l == 0) continue;                                           // uniform
        const uint32_t c = code[l]++;
        if (lane == 0) sorted[offs[l]] = (uint16_t)s;
        offs[l]++;
        const uint32_t r = inf_brev(c, l);
        const uint32_t e = inf_entry(kind, s, l);
        if (l <= bits) {
            for (int k = lane; k < (1 << (bits - l)); k += W) lut[r + ((uint32_t)k << l)] = e;
        } else if (sub_cap > 0) {
            const uint32_t root = lut[r & ((1u << bits) - 1u)];          // written before the last sync
            if (root & kInfESub) {
                const int      sb = (int)((root >> 8) & 31u), rest = l - bits;
                const uint32_t base = (root >> 16) + (r >> bits);
                for (int k = lane; k < (1 << (sb - rest)); k += W) lut[base + ((uint32_t)k << rest)] = e;
            }
        }
    }
    InfLanes::sync();
    return true;
}

// one symbol of a canonical code that the first-level table did not resolve (or any symbol): bit by bit
PSS_IHD_COLD int inf_slow(uint32_t bits, const uint16_t *count, const uint16_t *sorted, int &len_out)
{
    int      code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) { len_out = l; return sorted[index + (code - first)]; }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    len_out = 15;
    return -1;
}

// a symbol of a small "plain" code (the code-length code of a dynamic block)
PSS_IHD int inf_decode(InfBits &B, const uint32_t *lut, int bits, const uint16_t *count, const uint16_t *sorted)
{
    const uint32_t e = lut[B.peek(bits)];
    if (!(e & kInfESpecial)) {
        B.drop((int)(e & 31u));
        return (int)(e >> 16);
    }
    int       l;
    const int s = inf_slow((uint32_t)B.buf, count, sorted, l);
    B.drop(l);
    return s;
}

// A code longer than the first-level table: the canonical walk resumed behind the `bits` levels the table covers
// (their state was computed when the table was built), at most 15 - bits steps.  `word` = the next 32 bits of the stream
// (at least 15 of them valid).
PSS_IHD_COLD int inf_long(uint32_t word, const uint16_t *count, const uint16_t *sorted, uint32_t first0, uint32_t index0, int bits, int &len_out)
{
    int code = (int)(inf_brev(word & ((1u << bits) - 1u), bits) << 1);
    int      first = (int)first0, index = (int)index0;
    uint32_t rest = word >> bits;
    for (int l = bits + 1; l <= 15; l++) {
        code |= (int)(rest & 1u);
        rest >>= 1;
        const int c = count[l];
        if (code - c < first) { len_out = l; return sorted[index + (code - first)]; }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    len_out = 15;
    return -1;
}

// ---- the symbol loop ---------------------------------------------------------------------------------------------------
// Literal/length + distance symbols of one DEFLATE block up to its end-of-block code, warp uniform.  This is where the
// time goes: the loop is a chain of dependent steps (bit buffer -> table index -> shared-memory load -> code length ->
// bit buffer), and with tens of warps per SM it is bound by instruction issue -- so it is written for few instructions:
//   * one refill guarantees 32 valid bits; two literals, or a literal and a length with its extra bits, or a distance
//     with its extra bits are taken from those 32 bits with 32-bit shifts, and the 64-bit buffer is advanced once;
//   * the table entry is the decoded symbol (byte / base / bit counts): no second look-up, no arithmetic on symbols;
//   * while at least 260 bytes of output are left nothing checks the output bound (a step writes at most 1 + 258);
//     the last bytes of a block go through the same code with the checks compiled in;
//   * the first-level tables are read through their 32-bit shared-memory address, errors leave through one exit.
#if defined(__CUDA_ARCH__)
struct InfLut {
    uint32_t a;                                              // shared-space address of a uint32_t table
    __device__ __forceinline__ explicit InfLut(const uint32_t *p) : a((uint32_t)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const
    {
        uint32_t v;      // the address as one multiply-add (FMA pipe) instead of shift + add behind the index mask
        asm volatile("{\n\t.reg .u32 t;\n\tmad.lo.u32 t, %1, 4, %2;\n\tld.shared.u32 %0, [t];\n\t}" : "=r"(v) : "r"(i), "r"(a) : "memory");
        return v;
    }
};
#else
struct InfLut {
    const uint32_t *p;
    explicit InfLut(const uint32_t *q) : p(q) {}
    uint32_t operator[](uint32_t i) const { return p[i]; }
};
#endif

// one literal to the output, at p[OFF] (the pointer has been made opaque to the compiler, which would otherwise fall
// back to a generic store: say "global" explicitly).  Every lane stores the same byte to the same address -- the
// memory system makes one write of it -- which costs less than keeping a "lane 0" predicate alive through the loop.
template <int OFF = 0>
PSS_IHD void inf_store(uint8_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.u8 [%0+%2], %1;" ::"l"(p), "r"(v), "n"(OFF) : "memory");
#else
    p[OFF] = (uint8_t)v;
#endif
}

// LZ77 match: wp[0 .. len) = wp[-dist ..], by all lanes (byte i by lane i mod 32).  An overlapping match (dist < len)
// repeats its last `dist` bytes.  The general form, kept out of line: matches longer than 32 bytes.
PSS_IHD_COLD void inf_copy_long(uint8_t *wp, uint32_t dist, uint32_t len)
{
    const int lane = InfLanes::lane(), W = InfLanes::width();
    const uint8_t *src = wp - dist;
    for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)W) wp[i] = src[dist >= len ? i : i % dist];
}
// Matches of a BAM average nine bytes: up to 32 bytes are one load and one store per lane.  wpl = wp + lane.
// The bytes were written moments ago and come back from L2 (stores do not allocate in L1): a few hundred cycles that a
// load into a register would make the warp sit out at once (its consumer, or the next branch, waits on the
// scoreboard).  So the source is fetched ASYNCHRONOUSLY -- every lane cp.asyncs the aligned word around its own source
// byte into its own word of shared memory, tracked by the async-copy counter instead of a register scoreboard -- and the
// store is issued when the NEXT match (or the end of the loop) comes round, after the warp has decoded the symbols in
// between.  Nothing in between reads the bytes still owed: literals go to other addresses, and the next match settles
// the debt before it fetches.
struct InfPending {
    uint8_t *dst;                // where this lane's byte goes
    uint32_t sidx;               // ... and where it sits in the stage
    uint32_t owed;               // this lane has a byte to store
};
PSS_IHD void inf_settle(InfPending &P, uint32_t stage)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");         // this lane's own word has landed: no other lane's is needed
    if (P.owed) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(stage + P.sidx) : "memory");
        asm volatile("st.global.u8 [%0], %1;" ::"l"(P.dst), "r"(v) : "memory");
    }
    P.owed = 0u;
#else
    (void)P; (void)stage;
#endif
}
// `stage` = this LANE's word of the warp's stage
PSS_IHD void inf_copy(uint8_t *wp, uint8_t *wpl, uint32_t dist, uint32_t len, int lane, InfPending &P, uint32_t stage)
{
#if defined(__CUDA_ARCH__)
    inf_settle(P, stage);
    InfLanes::sync();                                        // the bytes written so far are visible to every lane
    if (len > 32u) { inf_copy_long(wp, dist, len); return; }                  // warp uniform
    uint32_t back = dist;
    if (dist < len) {                                                           // warp uniform: a run (binned qualities)
        // lane % dist without an integer division: (lane + 0.5) / dist is never within 1e-3 of an integer
        float r;                                             // 1 / dist, approximate (one MUFU): far more precise than needed
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)dist));
        const uint32_t q = (uint32_t)(((float)lane + 0.5f) * r);
        back = dist + q * dist;
    }
    const uintptr_t src = (uintptr_t)(wpl - back);           // this lane's source byte; it fetches the aligned word around it
    P.owed = (uint32_t)lane < len ? 1u : 0u;
    P.dst = wpl;
    P.sidx = (uint32_t)(src & 3u);
    if (P.owed) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(stage), "l"(src & ~(uintptr_t)3) : "memory");
#else
    (void)wpl; (void)lane; (void)P; (void)stage;
    const uint8_t *src = wp - dist;
    for (uint32_t i = 0; i < len; i++) wp[i] = src[dist >= len ? i : i % dist];
#endif
}

constexpr int      kInfMore = 101;                           // inf_loop<false>: fewer than kInfFastRoom bytes of output left
constexpr uint32_t kInfFastRoom = 260;                       // bytes of output a step without bound checks may need (1 + 258; 3 literals)

// The symbol loop proper.  A step takes two or three literals | [a literal and] one literal, end of block or match.
// CAREFUL = false: runs while op <= limit (= out_len - kInfFastRoom) and checks no output bound; returns kInfMore when
// it runs out of that guarantee.  CAREFUL = true: one symbol per step, every bound checked (the tail of a block).
// Returns kInfOk at the end-of-block code, or an error.
template <bool CAREFUL>
PSS_IHD int inf_loop(InfBits &B, InfLut lit, InfLut dst, InflateTables &T, uint8_t *out, uint32_t &op_io, uint32_t out_len, uint32_t limit)
{
    constexpr uint32_t kLitMask = (1u << kInfLitBits) - 1u, kDistMask = (1u << kInfDistBits) - 1u;
    const int  lane = InfLanes::lane();
    uint32_t   op = op_io;
#if defined(__CUDA_ARCH__)
    // opaque to the compiler: otherwise it re-derives the shared-memory addresses in every iteration
    asm volatile("" : "+r"(lit.a), "+r"(dst.a));
    asm volatile("" : "+l"(out));
#endif
    uint8_t *const outl = out + lane;
    uint32_t       bad_dist = 0;
    InfPending     pend;
    pend.owed = 0u; pend.dst = out; pend.sidx = 0;
#if defined(__CUDA_ARCH__)
    const uint32_t stage = (uint32_t)__cvta_generic_to_shared(T.stage + lane);       // this lane's word
#else
    const uint32_t stage = 0;
#endif
    int            rc;
    for (;;) {
        if (!CAREFUL && op > limit) { rc = kInfMore; break; }
        B.refill();
        uint32_t x = (uint32_t)B.buf;                        // 32 valid bits
        uint32_t e = lit[x & kLitMask];
        uint32_t used = 0;
        uint8_t *wp = out + op;
        if (!CAREFUL && (e & kInfELit)) {
            inf_store(wp, e >> 16);
            op++;
            used = e & 31u;                                  // a first-level hit: <= kInfLitBits bits
            x >>= used;                                      // >= 22 valid bits: any code and the extra bits of a length
            e = lit[x & kLitMask];
            if (e & kInfELit) {
                inf_store<1>(wp, e >> 16);
                op++;
                used += e & 31u;
                x >>= e & 31u;                               // >= 12 valid bits: enough for one more first-level literal
                e = lit[x & kLitMask];
                if (e & kInfELit) {                          // (anything else is looked up again after the next refill)
                    inf_store<2>(wp, e >> 16);
                    op++;
                    used += e & 31u;
                }
                B.drop((int)used);
                continue;
            }
            wp++;
        }
        if (e & (kInfESub | kInfESpecial | kInfELit)) {      // anything but a length with a first-level code
            if (e & kInfESub) e = lit[(e >> 16) + ((x >> kInfLitBits) & ((1u << ((e >> 8) & 31u)) - 1u))];
            if (e & kInfESpecial) {
                if ((e >> 16) == kInfSpWalk) {               // no table entry: the canonical walk decides
                    int       ll;
                    const int s = inf_long(x, T.lit_count, T.lit_sorted, T.lit_first, T.lit_index, kInfLitBits, ll);
                    if (s < 0) { rc = kInfBadSymbol; break; }
                    e = inf_entry(kInfKindLit, s, ll);
                }
                if (e & kInfESpecial) {
                    if ((e >> 16) != kInfSpEnd) { rc = kInfBadSymbol; break; }
                    B.drop((int)(used + (e & 31u)));
                    rc = kInfOk;
                    break;
                }
            }
            if (e & kInfELit) {
                if (CAREFUL && op >= out_len) { rc = kInfOutputOverrun; break; }
                inf_store(wp, e >> 16);
                op++;
                B.drop((int)(used + (e & 31u)));
                continue;
            }
        }
        // a match: length base + extra bits (code and extra bits: <= 20 bits of x), then the distance the same way (<= 28)
        const uint32_t tot = (e >> 8) & 31u;
        const uint32_t len = (e >> 16) + ((x & ((1u << tot) - 1u)) >> (e & 31u));
        B.drop((int)(used + tot));
        B.refill();
        x = (uint32_t)B.buf;
        uint32_t d = dst[x & kDistMask];
        if (d & kInfESpecial) {
            if ((d >> 16) != kInfSpWalk) { rc = kInfBadDistance; break; }
            int       ll;
            const int ds = inf_long(x, T.dist_count, T.dist_sorted, T.dist_first, T.dist_index, kInfDistBits, ll);
            if (ds < 0) { rc = kInfBadDistance; break; }
            d = inf_entry(kInfKindDist, ds, ll);
            if (d & kInfESpecial) { rc = kInfBadDistance; break; }
        }
        const uint32_t dtot = (d >> 8) & 31u;
        const uint32_t dist = (d >> 16) + ((x & ((1u << dtot) - 1u)) >> (d & 31u));
        B.drop((int)dtot);
        // a distance that reaches back before the block is an error; the fast loop only notes it and reports it when
        // it ends (the bytes such a match copies come from inside the buffer -- up to 32 KiB before the block, the carry
        // region lies there -- and the block is rejected anyway)
#if defined(__CUDA_ARCH__)
        if (CAREFUL) { if (dist > op) { rc = kInfBadDistance; break; } }
        else bad_dist |= (dist > op) ? 1u : 0u;
#else
        if (dist > op) { rc = kInfBadDistance; break; }       // (the host build has no such slack before its buffers)
#endif
        if (CAREFUL && len > out_len - op) { rc = kInfOutputOverrun; break; }
        inf_copy(wp, outl + op, dist, len, lane, pend, stage);
        op += len;
    }
    inf_settle(pend, stage);
    op_io = op;
    if (bad_dist) return kInfBadDistance;
    return rc;
}
PSS_IHD_COLD int inf_loop_careful(InfBits &B, InflateTables &T, uint8_t *out, uint32_t &op, uint32_t out_len)
{
    return inf_loop<true>(B, InfLut(T.lit_lut), InfLut(T.dist_lut), T, out, op, out_len, 0u);
}

PSS_IHD int inf_symbols(InfBits &B, InflateTables &T, uint8_t *out, uint32_t &op, uint32_t out_len)
{
    if (out_len >= kInfFastRoom) {
        const int rc = inf_loop<false>(B, InfLut(T.lit_lut), InfLut(T.dist_lut), T, out, op, out_len, out_len - kInfFastRoom);
        if (rc != kInfMore) return rc;
    }
    return inf_loop_careful(B, T, out, op, out_len);
}

// ---- one BGZF payload -------------------------------------------------------------------------------------------------
// Inflate the raw DEFLATE stream [in, in + in_len) into out[0 .. out_len); out_len is the ISIZE of the BGZF trailer
// and must be met exactly.  Returns kInfOk or an error code (warp uniform).  `out` may have any alignment.
PSS_IHD int inflate_block(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, InflateTables &T)
{
    const int lane = InfLanes::lane(), W = InfLanes::width();
    InfBits   B;
    B.open(in, in_len, T.window);
    uint32_t  mis = (uint32_t)((uintptr_t)in & 3u);
    uint32_t  op = 0;
    int       last;
    do {
        B.refill();
        last = (int)B.get(1);
        const int type = (int)B.get(2);
        if (type == 0) {
            // stored: skip to the byte boundary, LEN, NLEN, bytes
            B.drop(B.cnt & 7);
            B.refill();
            const uint32_t len = B.get(16);
            B.refill();
            const uint32_t nlen = B.get(16);
            if ((len ^ 0xffffu) != nlen) return kInfBadStored;
            if (op + len > out_len) return kInfOutputOverrun;
            // the bit buffer holds whole bytes now; the remaining ones come straight from memory
            const uint32_t at = B.bytes_used(mis);
            if (at + len > in_len) return kInfInputOverrun;
            for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)W) out[op + i] = in[at + i];
            op += len;
            // reopen the reader behind the stored bytes
            in += at + len;
            in_len -= at + len;
            mis = (uint32_t)((uintptr_t)in & 3u);
            B.open(in, in_len, T.window);
            InfLanes::sync();
            continue;
        }
        if (type == 3) return kInfBadBlockType;
        if (type == 1) {
            for (int s = lane; s < 288; s += W) T.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            for (int s = lane; s < 32; s += W) T.lens[288 + s] = 5;
            InfLanes::sync();
            if (!inf_build(T.lens, 288, T.lit_count, T.lit_sorted, T.lit_lut, kInfLitBits, kInfKindLit, kInfLitSub, &T.lit_first, &T.lit_index)) return kInfBadCodeLengths;
            if (!inf_build(T.lens + 288, 32, T.dist_count, T.dist_sorted, T.dist_lut, kInfDistBits, kInfKindDist, 0, &T.dist_first, &T.dist_index)) return kInfBadCodeLengths;
        } else {
            B.refill();
            const int hlit = (int)B.get(5) + 257, hdist = (int)B.get(5) + 1, hclen = (int)B.get(4) + 4;
            if (hlit > 286 || hdist > 30) return kInfBadCodeLengths;
            // code-length code: 19 symbols, 3 bits each, in the order of RFC 1951 3.2.7; its table reuses dist_lut /
            // dist_sorted / dist_count (7-bit codes at most), its lengths sit behind the real ones for a moment
            uint8_t *cl = T.lens + kInfMaxLit;           // 19 entries, overwritten by the distance lengths afterwards
            for (int i = lane; i < 19; i += W) cl[i] = 0;
            InfLanes::sync();
            for (int i = 0; i < hclen; i++) {
                B.refill();
                const uint32_t v = B.get(3);
                // order: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
                const int pos = i < 3 ? 16 + i : i == 3 ? 0 : (i & 1) ? (19 - i) / 2 : 8 + (i - 4) / 2;
                if (lane == 0) cl[pos] = (uint8_t)v;
            }
            InfLanes::sync();
            if (!inf_build(cl, 19, T.dist_count, T.dist_sorted, T.dist_lut, 7, kInfKindPlain, 0)) return kInfBadCodeLengths;
            // the hlit + hdist code lengths, run-length coded.  They are collected at lens[0 .. hlit + hdist) and the
            // distance part is moved to lens[288 ..) afterwards; the code-length table lives in dist_* meanwhile, and
            // cl[] may be overwritten once it is built.
            int n = 0, prev = 0;
            const int total = hlit + hdist;
            while (n < total) {
                B.refill();
                const int s = inf_decode(B, T.dist_lut, 7, T.dist_count, T.dist_sorted);
                if (s < 0) return kInfBadCodeLengths;
                int rep, val;
                if (s < 16) { rep = 1; val = s; prev = s; }
                else if (s == 16) { if (n == 0) return kInfBadCodeLengths; rep = 3 + (int)B.get(2); val = prev; }
                else if (s == 17) { rep = 3 + (int)B.get(3); val = 0; prev = 0; }
                else { rep = 11 + (int)B.get(7); val = 0; prev = 0; }
                if (n + rep > total) return kInfBadCodeLengths;
                // the collected lengths go to a staging area that cannot collide with cl[]: the lit_lut words (rebuilt
                // below anyway) hold them as bytes
                uint8_t *stage = reinterpret_cast<uint8_t *>(T.lit_lut);
                for (int i = lane; i < rep; i += W) stage[n + i] = (uint8_t)val;
                n += rep;
            }
            InfLanes::sync();
            {
                const uint8_t *stage = reinterpret_cast<const uint8_t *>(T.lit_lut);
                for (int s = lane; s < 288; s += W) T.lens[s] = s < hlit ? stage[s] : (uint8_t)0;
                for (int s = lane; s < 32; s += W) T.lens[288 + s] = s < hdist ? stage[hlit + s] : (uint8_t)0;
            }
            InfLanes::sync();
            if (T.lens[256] == 0) return kInfBadCodeLengths;             // no end-of-block code
            if (!inf_build(T.lens, 288, T.lit_count, T.lit_sorted, T.lit_lut, kInfLitBits, kInfKindLit, kInfLitSub, &T.lit_first, &T.lit_index)) return kInfBadCodeLengths;
            if (!inf_build(T.lens + 288, 32, T.dist_count, T.dist_sorted, T.dist_lut, kInfDistBits, kInfKindDist, 0, &T.dist_first, &T.dist_index)) return kInfBadCodeLengths;
        }
        // ---- the symbol loop (warp uniform)
        {
            const int rc = inf_symbols(B, T, out, op, out_len);
            if (rc != kInfOk) return rc;
        }
    } while (!last);
    if (B.bytes_used(mis) > in_len) return kInfInputOverrun;       // bits beyond the payload were consumed (they read as 0)
    if (op != out_len) return kInfSizeMismatch;
    return kInfOk;
}

}  // namespace pssgpu
