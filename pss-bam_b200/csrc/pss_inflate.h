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
// Tables (per warp, in shared memory; 3.7 KB): a 10-bit first-level table for literal/length codes and an 8-bit one for
// distance codes (entry = symbol << 4 | code length; 0 = longer code), and for the rare longer codes the canonical
// count/sorted-symbol arrays walked bit by bit.
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

constexpr int kInfLitBits  = 10;
constexpr int kInfDistBits = 8;
constexpr int kInfMaxLit   = 288;
constexpr int kInfMaxDist  = 32;

struct InflateTables {                       // one per warp
    uint16_t lit_lut[1 << kInfLitBits];      // symbol << 4 | length, 0: code longer than kInfLitBits
    uint16_t dist_lut[1 << kInfDistBits];
    uint16_t lit_sorted[kInfMaxLit];         // symbols in canonical order (by code length, then symbol)
    uint16_t dist_sorted[kInfMaxDist];
    uint16_t lit_count[16], dist_count[16];  // codes per length
    uint32_t lit_first, lit_index;           // state of the canonical walk after kInfLitBits levels (longer codes resume there)
    uint32_t dist_first, dist_index;         // ... after kInfDistBits levels
    uint8_t  lens[kInfMaxLit + kInfMaxDist]; // code lengths of the block being set up
};

enum : int {
    kInfOk = 0,
    kInfBadBlockType = 1, kInfBadStored = 2, kInfBadCodeLengths = 3, kInfBadSymbol = 4, kInfBadDistance = 5,
    kInfOutputOverrun = 6, kInfInputOverrun = 7, kInfSizeMismatch = 8
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
// lane l keeps word (32 * window + l) in a register and the next window is already in flight; `in_words` bounds the
// reads (words beyond it read as 0 and raise kInfInputOverrun only if their bits are consumed).
struct InfBits {
    const uint32_t *words;       // 4-byte aligned
    uint32_t        n_words;     // words that may be read
    uint32_t        next;        // index of the next word to enter the bit buffer
    uint64_t        buf;
    int             cnt;         // valid bits in buf
    uint32_t        win, win_next;   // device: this lane's word of the current / next 32-word window

    PSS_IHD uint32_t load(uint32_t i) const
    {
#if defined(__CUDA_ARCH__)
        return i < n_words ? __ldg(words + i) : 0u;
#else
        return i < n_words ? words[i] : 0u;
#endif
    }
    PSS_IHD void open(const uint8_t *p, uint32_t n_bytes)
    {
        const uintptr_t a = (uintptr_t)p;
        const uint32_t  mis = (uint32_t)(a & 3u);
        words = reinterpret_cast<const uint32_t *>(a - mis);
        n_words = (n_bytes + mis + 3u) >> 2;
        next = 0;
#if defined(__CUDA_ARCH__)
        win = load((uint32_t)InfLanes::lane());
        win_next = load(32u + (uint32_t)InfLanes::lane());
#else
        win = win_next = 0;
#endif
        buf = 0; cnt = 0;
        const uint32_t w = take_word();
        buf = (uint64_t)(w >> (8u * mis));
        cnt = 32 - 8 * (int)mis;
    }
    PSS_IHD uint32_t take_word()
    {
#if defined(__CUDA_ARCH__)
        const uint32_t w = InfLanes::bcast(win, (int)(next & 31u));
        next++;
        if ((next & 31u) == 0u) {                  // warp uniform
            win = win_next;
            win_next = load(next + 32u + (uint32_t)InfLanes::lane());
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

// ---- table set-up ---------------------------------------------------------------------------------------------------
// Canonical Huffman code of `n` symbols with code lengths lens[0..n): fills count[], sorted[] and the first-level
// table lut (2^bits entries).  Returns false for an over-subscribed or (non-trivially) incomplete code.  All lanes
// call it together; the symbol loop is uniform, the table fill is spread over the lanes.
PSS_IHD_COLD bool inf_build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sorted, uint16_t *lut, int bits, bool flag_literals = false,
                            uint32_t *resume_first = nullptr, uint32_t *resume_index = nullptr)
{
    const int lane = InfLanes::lane(), W = InfLanes::width();
    for (int i = lane; i < 16; i += W) count[i] = 0;
    for (int i = lane; i < (1 << bits); i += W) lut[i] = 0;
    InfLanes::sync();
    if (lane == 0)
        for (int s = 0; s < n; s++) count[lens[s]]++;
    InfLanes::sync();
    // offsets / first codes per length (uniform, in registers)
    uint32_t offs[16], code[16];
    int      left = 1;
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
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (l == 0) continue;                                           // uniform
        const uint32_t c = code[l]++;
        if (lane == 0) sorted[offs[l]] = (uint16_t)s;
        offs[l]++;
        if (l <= bits) {
#if defined(__CUDA_ARCH__)
            const uint32_t r = __brev(c) >> (32 - l);
#else
            uint32_t r = 0;
            for (int b = 0; b < l; b++) r |= ((c >> b) & 1u) << (l - 1 - b);
#endif
            // literal/length table: bit 15 marks a literal, so that the symbol loop decides "literal, resolved" with one test
            const uint16_t e = (uint16_t)((s << 4) | l | ((flag_literals && s < 256) ? 0x8000 : 0));
            for (int k = lane; k < (1 << (bits - l)); k += W) lut[r + ((uint32_t)k << l)] = e;
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

PSS_IHD int inf_decode(InfBits &B, const uint16_t *lut, int bits, const uint16_t *count, const uint16_t *sorted)
{
    const uint32_t e = lut[B.peek(bits)];
    if (e & 15u) {
        B.drop((int)(e & 15u));
        return (int)(e >> 4);
    }
    int       l;
    const int s = inf_slow((uint32_t)B.buf, count, sorted, l);
    B.drop(l);
    return s;
}

// A code longer than the first-level table: the canonical walk resumed behind the `bits` levels the table covers
// (their state was computed when the table was built), at most 15 - bits steps.  `word` = the next 32 bits of the stream.
PSS_IHD int inf_long(uint32_t word, const uint16_t *count, const uint16_t *sorted, uint32_t first0, uint32_t index0, int bits, int &len_out)
{
#if defined(__CUDA_ARCH__)
    int code = (int)((__brev(word) >> (32 - bits)) << 1);
#else
    uint32_t rv = 0;
    for (int b = 0; b < bits; b++) rv |= ((word >> b) & 1u) << (bits - 1 - b);
    int code = (int)(rv << 1);
#endif
    int      first = (int)first0, index = (int)index0;
    uint32_t rest = word >> bits;
#pragma unroll
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

// length / distance bases and extra bits (RFC 1951 3.2.5), packed: base | extra << 16
PSS_IHD uint32_t inf_len_code(int s)        // s = symbol - 257, 0..28
{
    // extra bits: 0 for s < 8, then (s - 4) / 4; code 28 is the literal 258
    const uint32_t eb = s < 8 ? 0u : (uint32_t)((s - 4) >> 2);
    const uint32_t base = s < 8 ? (uint32_t)(3 + s) : (uint32_t)(3 + ((4 + (s & 3)) << eb));
    return s == 28 ? 258u : (base | (eb << 16));
}
PSS_IHD uint32_t inf_dist_code(int s)       // 0..29
{
    const uint32_t eb = s < 4 ? 0u : (uint32_t)((s - 2) >> 1);
    const uint32_t base = s < 4 ? (uint32_t)(1 + s) : (uint32_t)(1 + ((2 + (s & 1)) << eb));
    return base | (eb << 16);
}

// ---- the symbol loop ---------------------------------------------------------------------------------------------------
// Literal/length + distance symbols of one DEFLATE block up to its end-of-block code, warp uniform.  This is where the
// time goes: the loop is a chain of dependent steps (bit buffer -> table index -> shared-memory load -> code length ->
// bit buffer), and with tens of warps per SM it is bound by instruction issue -- so the literal path is kept to a
// handful of instructions: the first-level tables are read through their 32-bit shared-memory address (no generic
// pointer arithmetic in the loop), the output position is one register, errors leave through one exit.
#if defined(__CUDA_ARCH__)
struct InfLut {
    uint32_t a;                                              // shared-space address of a uint16_t table
    __device__ __forceinline__ explicit InfLut(const uint16_t *p) : a((uint32_t)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a + 2u * i) : "memory");
        return v;
    }
};
#else
struct InfLut {
    const uint16_t *p;
    explicit InfLut(const uint16_t *q) : p(q) {}
    uint32_t operator[](uint32_t i) const { return p[i]; }
};
#endif

// one literal to the output (the pointer has been made opaque to the compiler, which would otherwise fall back to a
// generic store: say "global" explicitly)
PSS_IHD void inf_store(uint8_t *p, uint8_t v, bool writer)
{
#if defined(__CUDA_ARCH__)
    if (writer) asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"((uint32_t)v) : "memory");
#else
    if (writer) *p = v;
#endif
}

// LZ77 match: wp[0 .. len) = wp[-dist ..], by all lanes (byte i by lane i mod 32).  An overlapping match (dist < len)
// repeats its last `dist` bytes.  Matches of a BAM average nine bytes: one predicated load/store pair.
PSS_IHD void inf_copy(uint8_t *wp, uint32_t dist, uint32_t len, int lane, int W)
{
    const uint8_t *src = wp - dist;
#if defined(__CUDA_ARCH__)
    (void)W;
    for (uint32_t i = (uint32_t)lane; i < len; i += 32u) {
        const uint32_t k = dist >= len ? i : i % dist;
        uint32_t v;
        asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(src + k) : "memory");
        asm volatile("st.global.u8 [%0], %1;" ::"l"(wp + i), "r"(v) : "memory");
    }
#else
    for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)W) wp[i] = src[dist >= len ? i : i % dist];
#endif
}

PSS_IHD int inf_symbols(InfBits &B, InflateTables &T, uint8_t *out, uint32_t &op_io, uint32_t out_len)
{
    const int lane = InfLanes::lane(), W = InfLanes::width();
    InfLut    lit(T.lit_lut), dst(T.dist_lut);
    uint8_t  *wp = out + op_io;                              // next output byte
    uint32_t  room = out_len - op_io;                        // bytes that may still be written
#if defined(__CUDA_ARCH__)
    // opaque to the compiler: otherwise it re-derives the shared-memory addresses (six instructions) in every iteration
    asm volatile("" : "+r"(lit.a), "+r"(dst.a));
    asm volatile("" : "+l"(wp));
#endif
    const bool writer = lane == 0;
    int        rc = kInfOk;
    for (;;) {
        // ---- literals: bit buffer -> table -> literal flag, nothing else
        uint32_t e;
        for (;;) {
            B.refill();                                      // >= 33 bits: two literal codes (<= 15 bits each) fit
            e = lit[B.peek(kInfLitBits)];
            if (!(e & 0x8000u) || room == 0u) break;
            B.drop((int)(e & 15u));
            inf_store(wp, (uint8_t)(e >> 4), writer);
            wp++;
            room--;
            e = lit[B.peek(kInfLitBits)];                    // the second one on the same fill
            if (!(e & 0x8000u) || room == 0u) break;
            B.drop((int)(e & 15u));
            inf_store(wp, (uint8_t)(e >> 4), writer);
            wp++;
            room--;
        }
        // ---- everything else: longer codes, lengths, end of block, errors
        B.refill();                                          // (the entry `e` came from the low bits, which a fill leaves alone)
        uint32_t l = e & 15u;
        int      s = (int)((e & 0x7fffu) >> 4);
        if (l == 0u) {                                       // a code longer than the first-level table (7 % of the symbols of a BAM)
            int ll;
            s = inf_long((uint32_t)B.buf, T.lit_count, T.lit_sorted, T.lit_first, T.lit_index, kInfLitBits, ll);
            l = (uint32_t)ll;
            if (s < 0) { rc = kInfBadSymbol; break; }
        }
        B.drop((int)l);
        if (s < 256) {
            if (room == 0u) { rc = kInfOutputOverrun; break; }
            inf_store(wp, (uint8_t)s, writer);
            wp++;
            room--;
            continue;
        }
        if (s == 256) break;
        if (s > 285) { rc = kInfBadSymbol; break; }
        const uint32_t lc = inf_len_tab(s - 257);
        const uint32_t len = (lc & 0xffffu) + B.get((int)(lc >> 16));
        B.refill();
        e = dst[B.peek(kInfDistBits)];
        l = e & 15u;
        int ds = (int)(e >> 4);
        if (l == 0u) {
            int ll;
            ds = inf_long((uint32_t)B.buf, T.dist_count, T.dist_sorted, T.dist_first, T.dist_index, kInfDistBits, ll);
            l = (uint32_t)ll;
        }
        B.drop((int)l);
        if (ds < 0 || ds > 29) { rc = kInfBadDistance; break; }
        const uint32_t dc = inf_dist_tab(ds);
        const uint32_t dist = (dc & 0xffffu) + B.get((int)(dc >> 16));
        if (dist > out_len - room) { rc = kInfBadDistance; break; }
        if (len > room) { rc = kInfOutputOverrun; break; }
        InfLanes::sync();                                    // the bytes written so far are visible to every lane
        inf_copy(wp, dist, len, lane, W);
        wp += len;
        room -= len;
    }
    op_io = out_len - room;
    return rc;
}

// ---- one BGZF payload -------------------------------------------------------------------------------------------------
// Inflate the raw DEFLATE stream [in, in + in_len) into out[0 .. out_len); out_len is the ISIZE of the BGZF trailer
// and must be met exactly.  Returns kInfOk or an error code (warp uniform).  `out` may have any alignment.
PSS_IHD int inflate_block(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, InflateTables &T)
{
    const int lane = InfLanes::lane(), W = InfLanes::width();
    InfBits   B;
    B.open(in, in_len);
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
            B.open(in, in_len);
            InfLanes::sync();
            continue;
        }
        if (type == 3) return kInfBadBlockType;
        if (type == 1) {
            for (int s = lane; s < 288; s += W) T.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            for (int s = lane; s < 32; s += W) T.lens[288 + s] = 5;
            InfLanes::sync();
            if (!inf_build(T.lens, 288, T.lit_count, T.lit_sorted, T.lit_lut, kInfLitBits, true, &T.lit_first, &T.lit_index)) return kInfBadCodeLengths;
            if (!inf_build(T.lens + 288, 32, T.dist_count, T.dist_sorted, T.dist_lut, kInfDistBits, false, &T.dist_first, &T.dist_index)) return kInfBadCodeLengths;
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
            if (!inf_build(cl, 19, T.dist_count, T.dist_sorted, T.dist_lut, 7)) return kInfBadCodeLengths;
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
            if (!inf_build(T.lens, 288, T.lit_count, T.lit_sorted, T.lit_lut, kInfLitBits, true, &T.lit_first, &T.lit_index)) return kInfBadCodeLengths;
            if (!inf_build(T.lens + 288, 32, T.dist_count, T.dist_sorted, T.dist_lut, kInfDistBits, false, &T.dist_first, &T.dist_index)) return kInfBadCodeLengths;
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
