// pss_crc32.h -- CRC-32 (the gzip polynomial) of an inflated BGZF block, computed by the warp that inflated it, and a
// host build of the same lane code that the CPU test-suite checks against zlib (tests/host_emul).  The host build is a
// TEST of this logic; the product library only ever executes it inside the inflate kernel.
//
// Why: every BGZF block ends in the CRC32 and ISIZE of its payload (SAM spec 4.1, RFC 1952 2.3.1).  htslib -- the
// `samtools view` child the reference reads its input from (pss-bam.c:148-162, fragkon.c:84-93) -- verifies both and
// fails on a mismatch, so a damaged file never reaches the reference's tally.  Same here: a block whose CRC is wrong
// fails the feed.
//
// How.  A CRC is linear over GF(2): the register after a message is the XOR of what every 32-bit word of it contributes,
// each advanced (multiplied by a power of x modulo the polynomial) by the bytes that follow it.  The decode is serial,
// the check need not be: lane l of the warp takes the words l, l + 32, l + 64, ... of the block -- so a warp load is one
// coalesced 128-byte row -- and keeps
//     u  <-  advance(u, 128 bytes) ^ word                (four look-ups in 256-entry tables, like slicing-by-4)
// After its last word, lane l owes an advance by the 4 .. 128 bytes that follow that word: one carry-less
// multiplication modulo the polynomial with x^(32 k), k = 1 .. 32; the XOR of the 32 results is the register.  The bytes
// before the first aligned word and behind the last whole word go through the ordinary byte table.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PSS_CHD __host__ __device__ __forceinline__
#else
#define PSS_CHD inline
#endif

namespace pssgpu {

constexpr uint32_t kCrcPoly   = 0xedb88320u;                 // reflected: bit 31 is x^0, a right shift multiplies by x
constexpr int      kCrcLanes  = 32;                          // words per row = lanes that share a block
constexpr int      kCrcT0     = 0;                           // table layout (32-bit words): the byte table ...
constexpr int      kCrcAdv    = 256;                         // ... four tables "advance byte k of the register by one row" ...
constexpr int      kCrcXk     = 256 + 4 * 256;               // ... x^(32 k) for k = 0 .. 32
constexpr int      kCrcSmemWords = kCrcXk;                   // what the lane loop looks up (kept in shared memory)
constexpr int      kCrcTableWords = kCrcXk + 33;

// a * b modulo the polynomial (zlib's multmodp, branch-free)
PSS_CHD uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
    for (int i = 0; i < 32; i++) {
        p ^= b & (0u - ((a >> (31 - i)) & 1u));
        b = (b >> 1) ^ (kCrcPoly & (0u - (b & 1u)));
    }
    return p;
}

// host: fill tab[0 .. kCrcTableWords)
inline void crc32_build_tables(uint32_t *tab)
{
    for (uint32_t b = 0; b < 256; b++) {                     // the register `b` advanced by one byte
        uint32_t c = b;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (kCrcPoly & (0u - (c & 1u)));
        tab[kCrcT0 + b] = c;
    }
    uint32_t x32 = 0x80000000u;                              // x^0
    for (int k = 0; k < 32; k++) x32 = (x32 >> 1) ^ (kCrcPoly & (0u - (x32 & 1u)));      // x^32
    tab[kCrcXk] = 0x80000000u;
    for (int k = 1; k <= 32; k++) tab[kCrcXk + k] = crc_mulmod(tab[kCrcXk + k - 1], x32);
    const uint32_t xrow = tab[kCrcXk + kCrcLanes];           // x^(32 * 32): one row of 32 words
    for (int k = 0; k < 4; k++)
        for (uint32_t b = 0; b < 256; b++) tab[kCrcAdv + 256 * k + b] = crc_mulmod(xrow, b << (8 * k));
}

PSS_CHD uint32_t crc_byte(const uint32_t *t0, uint32_t s, uint32_t byte) { return t0[(s ^ byte) & 0xffu] ^ (s >> 8); }

PSS_CHD uint32_t crc_load_word(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldcg(p);                                        // written a moment ago by this warp: read where the stores went (L2)
#else
    return *p;
#endif
}
PSS_CHD uint32_t crc_load_byte(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldcg(p);
#else
    return *p;
#endif
}

// What lane `lane` contributes to the register after the n_words aligned words at q: its partial u and, in *shift, the
// number of words k (1 .. 32) its last word is followed by (0: the lane has no word).  first_xor is folded into word 0
// (the register the bytes before q left behind).  t = the tables (shared memory on the device).
PSS_CHD uint32_t crc32_lane_partial(const uint32_t *q, uint32_t n_words, uint32_t lane, uint32_t first_xor, const uint32_t *t, uint32_t *shift)
{
    uint32_t idx = lane;
    if (idx >= n_words) { *shift = 0; return 0u; }
    uint32_t u = crc_load_word(q + idx) ^ (lane == 0 ? first_xor : 0u);
    const uint32_t *a = t + kCrcAdv;
    for (idx += kCrcLanes; idx < n_words; idx += kCrcLanes)
        u = a[u & 0xffu] ^ a[256 + ((u >> 8) & 0xffu)] ^ a[512 + ((u >> 16) & 0xffu)] ^ a[768 + (u >> 24)] ^ crc_load_word(q + idx);
    *shift = n_words - (idx - kCrcLanes);
    return u;
}

// the bytes before the first aligned word of [p, p + n): returns the register after them, *head = how many
PSS_CHD uint32_t crc32_head(const uint8_t *p, uint32_t n, const uint32_t *t, uint32_t *head)
{
    uint32_t h = (4u - (uint32_t)((uintptr_t)p & 3u)) & 3u;
    if (h > n) h = n;
    uint32_t s = 0xffffffffu;
    for (uint32_t i = 0; i < h; i++) s = crc_byte(t + kCrcT0, s, crc_load_byte(p + i));
    *head = h;
    return s;
}
// the bytes behind the last whole word, and the final inversion
PSS_CHD uint32_t crc32_tail(const uint8_t *p, uint32_t n_tail, uint32_t s, const uint32_t *t)
{
    for (uint32_t i = 0; i < n_tail; i++) s = crc_byte(t + kCrcT0, s, crc_load_byte(p + i));
    return ~s;
}

#if defined(__CUDACC__)
// CRC-32 of [p, p + n), by the 32 lanes of a warp (all must call; the result is warp uniform).  t = the first
// kCrcSmemWords table words in shared memory, xk = the x^(32 k) table (global memory).
__device__ __forceinline__ uint32_t crc32_warp(const uint8_t *p, uint32_t n, const uint32_t *t, const uint32_t *__restrict__ xk)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t       head;
    uint32_t       s = crc32_head(p, n, t, &head);                            // <= 3 bytes, every lane the same
    const uint32_t n_words = (n - head) >> 2;
    if (n_words) {
        uint32_t       shift;
        const uint32_t u = crc32_lane_partial(reinterpret_cast<const uint32_t *>(p + head), n_words, lane, s, t, &shift);
        uint32_t       c = shift ? crc_mulmod(__ldg(xk + shift), u) : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, d);
        s = c;
    }
    return crc32_tail(p + head + 4u * n_words, (n - head) & 3u, s, t);
}
#endif

}  // namespace pssgpu
