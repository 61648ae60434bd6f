// pss_kernels.cuh -- the sm_100a kernels of the pss-bam hot path.
//
//   pack_kernel      K6  ASCII contig -> 4-bit packed genome groups
//   spectrum_kernel  K5  genome-kmer-count.c:68-79 + kmer.c:43-110 as a flat 4^k histogram
//   tally_kernel     K1-K4  SAM text tile -> records -> filters -> genome gather
//                           -> pss-bam count tables / fragkon end-context histograms
//
// Nothing here is GEMM shaped; the kernels are byte scans, gathers and
// histograms bound by HBM and by integer issue rate.  The per-record logic is
// in pss_record.h (shared with the CPU test build).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pss_record.h"

namespace pssgpu {

// ---------------------------------------------------------------------------
// PTX: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, size multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ===========================================================================
// K6  genome pack
// ===========================================================================
struct PackArgs {
    const uint8_t *src;        // ASCII of bases [0, n_bases) of this piece (device)
    uint64_t       n_bases;
    uint64_t      *dst;        // first group of the piece
    uint64_t       gbase;      // global base index of src[0] (multiple of 16)
    uint64_t      *exc_pos;    // "other" symbol log
    uint8_t       *exc_chr;
    unsigned long long *exc_n;
    uint64_t       exc_cap;
    uint32_t      *flags;      // bit 0: NUL byte seen
};

__global__ void __launch_bounds__(256) pack_kernel(PackArgs A)
{
    const uint64_t n_groups = (A.n_bases + 15) / 16;
    const bool     aligned = ((uintptr_t)A.src & 15u) == 0;
    for (uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < n_groups;
         gi += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b0 = gi * 16;
        const int      nv = (A.n_bases - b0 >= 16) ? 16 : (int)(A.n_bases - b0);
        uint32_t       om = 0, nul = 0;
        uint64_t       grp;
        if (aligned && nv == 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(A.src + b0));
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
            grp = pack_group([&](int j) { return (uint8_t)(w[j >> 2] >> (8 * (j & 3))); }, 16, &om, &nul);
        } else {
            const uint8_t *p = A.src + b0;
            grp = pack_group([&](int j) { return p[j]; }, nv, &om, &nul);
        }
        A.dst[gi] = grp;
        if (nul) atomicOr(A.flags, 1u);
        while (om) {
            const int j = __ffs((int)om) - 1;
            om &= om - 1;
            const unsigned long long slot = atomicAdd(A.exc_n, 1ull);
            if (slot < A.exc_cap) {
                A.exc_pos[slot] = A.gbase + b0 + (uint64_t)j;
                A.exc_chr[slot] = upper_c(A.src[b0 + j]);
            }
        }
    }
}

// ===========================================================================
// K5  genome k-mer spectrum
// ===========================================================================
// One thread per packed group: the 16 k-mers that START in the group (they
// reach at most 13 bases into the next one).  K-mers touching a non-ACGT
// symbol -- which includes the padding between contigs, so no k-mer spans two
// contigs (genome-kmer-count.c:56-58) -- are skipped (kmer.c:94-96,207-208).
constexpr int kSpectrumSmemK = 6;     // 4^6 u32 bins = 16 KB privatised per CTA

// CT = counter type of the global table: unsigned int while the genome has fewer than 2^32 positions (no bin can
// overflow; 4^12 bins are then 64 MiB and stay in L2), unsigned long long otherwise.
template <bool SMEM, typename CT>
__global__ void __launch_bounds__(256) spectrum_kernel(const uint64_t *__restrict__ groups, uint64_t g_begin, uint64_t g_end,
                                                       int K, CT *__restrict__ counts)
{
    __shared__ uint32_t s_hist[SMEM ? (1 << (2 * kSpectrumSmemK)) : 1];
    const uint32_t n_bins = 1u << (2 * K);
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const uint32_t kmask = n_bins - 1u;
    const uint64_t kmask64 = kmask;
    for (uint64_t gi = g_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < g_end;
         gi += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g0 = __ldg(groups + gi), g1 = __ldg(groups + gi + 1);
        const uint64_t codes = (uint64_t)(uint32_t)g0 | ((uint64_t)(uint32_t)g1 << 32);
        const uint64_t cls = (g0 >> 32) | (g1 & 0xffffffff00000000ull);
        const uint64_t bad = (cls | (cls >> 1)) & kEvenBits;
        if ((uint32_t)bad == 0x55555555u) continue;          // whole group invalid (padding / N run)
#pragma unroll
        for (int o = 0; o < 16; o++) {
            if (((bad >> (2 * o)) & kmask64) != 0) continue;
            const uint32_t w = (uint32_t)(codes >> (2 * o)) & kmask;
            const uint32_t idx = rev_fields32(w) >> (32 - 2 * K);
            if (SMEM) atomicAdd(&s_hist[idx], 1u);
            else      atomicAdd(counts + idx, (CT)1);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(counts + i, (CT)v);
        }
    }
}
// u32 spectrum -> u64 output
__global__ void __launch_bounds__(256) widen_kernel(const unsigned int *__restrict__ in, unsigned long long *__restrict__ out, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// ===========================================================================
// K1-K4  SAM tile scan + per-record tally
// ===========================================================================
// tuning knobs (defaults = the measured best; other values are only used by tuning builds)
#ifndef PSS_TALLY_CTAS_PER_SM
#define PSS_TALLY_CTAS_PER_SM 4
#endif
#ifndef PSS_TILE_MAIN
#define PSS_TILE_MAIN 32768
#endif
#ifndef PSS_TALLY_THREADS
#define PSS_TALLY_THREADS 256
#endif
constexpr int kTileMain   = PSS_TILE_MAIN;                            // bytes of SAM a CTA owns per tile
constexpr int kTileOver   = 2048;                             // look-ahead so the last owned record is whole
constexpr int kPrefix     = 16;                               // bytes before the tile (is byte -1 a '\n'?)
constexpr int kTileSpan   = kPrefix + kTileMain + kTileOver;  // 34832 staged bytes
constexpr int kThreads    = PSS_TALLY_THREADS;
constexpr int kWarps      = kThreads / 32;
constexpr int kWords      = (kTileSpan + 31) / 32;                // 32-byte chunks = mask words (1089)
constexpr int kWordsPerThread = (kWords + kThreads - 1) / kThreads;   // 5
constexpr int kRecCap     = 1024;                             // records materialised per pass
constexpr int kCacheContigs = 64;                             // contig table kept in shared memory when it fits
constexpr int kCacheNames   = 1024;
constexpr int kCacheSlots   = 128;
constexpr int kFlushEvery   = 1900;                           // warp iterations between flushes of the 16-bit partial sums

struct ContigCache {
    uint32_t n;                                               // 0: not cached, use the global table
    uint32_t hash[kCacheContigs];
    uint16_t name_w0[kCacheContigs], name_len[kCacheContigs]; // first word of the name in names32 / length in bytes
    uint64_t base_off[kCacheContigs], len[kCacheContigs];
    uint8_t  slot[kCacheSlots];                               // open addressing, 0xff = empty
    uint32_t names32[kCacheNames / 4 + kCacheContigs];        // names as zero-padded little-endian words
};

struct TallyShared {                                          // per-CTA tables, outcome counters, contig table
    uint32_t table[2 * 32 * 16];                              // CTA count tables [fwd|rev][row][cell]
    uint32_t stats[8];                                        // outcomes (pss-bam's in the fused mode)
    uint32_t stats_fk[8];                                     // fragkon's outcomes in the fused mode
    ContigCache cc;
};

struct TallySmem {
    alignas(128) uint8_t bytes[kWords * 32];                  // staged SAM text
    uint32_t le[kWords + 8];                                  // bit i of word w: byte 32w+i is <= 0x20
    uint32_t nl[kWords + 8];                                  //                  byte 32w+i is '\n'
    uint16_t wpre[kWords + 8];                                // newlines before word w
    uint16_t nlpos[kRecCap + 4];                              // newline positions of this pass
    TallyShared sh;
    uint32_t warp_sum[kWarps];
    uint32_t n_newlines;
    alignas(8) uint64_t bar;
};

struct TallyArgs {
    const uint8_t *sam;          // device, 16-byte aligned
    uint64_t       len;
    uint64_t       stream_off;   // offset of sam[0] within everything fed (debug log only)
    DevGenome      g;
    TallyCfg       cfg;          // pss-bam options (fragkon's when only fragkon runs)
    TallyCfg       cfg_fk;       // fragkon options of the fused mode
    uint32_t       names_bytes;  // total bytes of contig names
    unsigned long long *pss_tables;   // 2*(R+2)*16 u64: fwd then rev
    unsigned long long *fk_hist;      // 2*4^K u64: 5' then 3'
    unsigned long long *stats;        // kStN
    unsigned long long *stats_fk;     // kStN, fused mode only
    uint64_t           *dbg_off;      // debug log (may be null)
    int8_t             *dbg_code;
    unsigned long long *dbg_n;
    uint64_t            dbg_cap;
};

struct SmemAt {                  // absolute positions in the staged tile (4-byte aligned base)
    static constexpr int kLookBack = kPrefix;       // every record starts at >= kPrefix: 12 bytes back is always inside
    const uint8_t *p;
    __device__ __forceinline__ int lo() const { return 0; }
    __device__ __forceinline__ uint32_t operator()(int i) const { return p[i]; }
    __device__ __forceinline__ uint32_t word(int i) const { return *reinterpret_cast<const uint32_t *>(p + i); }
};
struct SmemRel {                 // positions relative to a record start (any alignment); records start at >= kPrefix
    static constexpr int kLookBack = kPrefix;
    const uint8_t *p;
    __device__ __forceinline__ int lo() const { return -kPrefix; }
    __device__ __forceinline__ uint32_t operator()(int i) const { return p[i]; }
    __device__ __forceinline__ uint32_t word(int i) const
    {
        return (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
    }
};
struct GlobalAt {                // positions relative to p; nothing before p is touched
    static constexpr int kLookBack = 0;
    const uint8_t *p;
    __device__ __forceinline__ int lo() const { return 0; }
    __device__ __forceinline__ uint32_t operator()(int i) const { return __ldg(p + i); }
    __device__ __forceinline__ uint32_t word(int i) const
    {
        return (uint32_t)__ldg(p + i) | ((uint32_t)__ldg(p + i + 1) << 8) | ((uint32_t)__ldg(p + i + 2) << 16) |
               ((uint32_t)__ldg(p + i + 3) << 24);
    }
};

// 0x80 in every byte of w that is <= 0x20 / that is '\n'
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &zle, uint32_t &znl)
{
    zle = ~(((w & 0x7f7f7f7fu) + 0x5f5f5f5fu) | w) & 0x80808080u;
    // among the flagged bytes (all <= 0x20, so bits 6 and 7 are clear) '\n' is the one whose low six bits equal 0x0a
    const uint32_t y = (w ^ 0x0a0a0a0au) & 0x3f3f3f3fu;
    znl = zle & ~(y + 0x7f7f7f7fu);
}
// 8 flag bytes (0x80 / 0) in two words -> 8-bit mask << 7, via two byte dot products
__device__ __forceinline__ uint32_t gather8(uint32_t z0, uint32_t z1)
{
    return __dp4a(z1, 0x80402010u, __dp4a(z0, 0x08040201u, 0u));
}

__device__ __forceinline__ void log_outcome(const TallyArgs &A, uint64_t goff, int code)
{
    if (A.dbg_n) {
        const unsigned long long slot = atomicAdd(A.dbg_n, 1ull);
        if (slot < A.dbg_cap) {
            A.dbg_off[slot] = A.stream_off + goff;
            A.dbg_code[slot] = (int8_t)code;
        }
    }
}
__device__ __forceinline__ int stat_slot(int code)
{
    return code == kCounted ? kStCounted : code == kNoContig ? kStNoContig : code == kFiltered ? kStFiltered
         : code == kParseFail ? kStParseFail : kStUndefined;
}

// k-th 4-byte word of the field [off, off+len), zero padded at the end
template <class B>
__device__ __forceinline__ uint32_t field_word(const B &b, int off, int len, int k)
{
    uint32_t  w = word_at(b, off + 4 * k);
    const int rem = len - 4 * k;
    if (rem < 4) w &= rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u);
    return w;
}
template <class B>
__device__ __forceinline__ uint32_t name_hash_words(const B &b, int off, int len)
{
    // every name hashes at least its first two (zero padded) words: no loop for names of up to 8 bytes
    uint32_t  h = kNameHashSeed ^ (uint32_t)len;
    const int nw = (len + 3) >> 2;
    h = (h ^ field_word(b, off, len, 0)) * 0x9E3779B1u;  h ^= h >> 15;
    h = (h ^ field_word(b, off, len, 1)) * 0x9E3779B1u;  h ^= h >> 15;
    for (int k = 2; k < nw; k++) {
        h = (h ^ field_word(b, off, len, k)) * 0x9E3779B1u;
        h ^= h >> 15;
    }
    return h;
}
// find_seq (fasta-genome-io.c:202-213) against the shared-memory copy of the
// contig table: hash and comparison run on whole words of RNAME.  The first
// two probes are straight-line code (with 128 slots for <= 64 names a third
// probe is rare), so the lanes of a warp do not drift apart here.
template <class B>
__device__ __forceinline__ bool cache_name_equal(const ContigCache &C, uint32_t v, const B &b, int off, int len,
                                                  uint32_t w0, uint32_t w1)
{
    const uint32_t *nm = C.names32 + C.name_w0[v];
    bool eq = (int)C.name_len[v] == len;
    if (len > 0) eq = eq && (nm[0] == w0);
    if (len > 4) eq = eq && (nm[1] == w1);
    if (len > 8) {                                    // long names: the remaining words
        const int nw = (len + 3) >> 2;
        for (int k = 2; eq && k < nw; k++) eq = (nm[k] == field_word(b, off, len, k));
    }
    return eq;
}
template <class B>
__device__ __forceinline__ int cache_find(const ContigCache &C, const B &b, int off, int len, uint64_t &base, uint64_t &clen)
{
    const uint32_t h = name_hash_words(b, off, len);
    const uint32_t w0 = field_word(b, off, len, 0), w1 = field_word(b, off, len, 1);
    uint32_t s = h & (kCacheSlots - 1);
    int      found = -1;
    bool     open = true;                             // still probing
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const uint32_t v = C.slot[s];
        const bool occupied = (v != 0xffu);
        const uint32_t vv = occupied ? v : 0u;
        const bool hit = open && occupied && C.hash[vv] == h && cache_name_equal(C, vv, b, off, len, w0, w1);
        if (hit) found = (int)vv;
        open = open && occupied && !hit;
        s = (s + 1) & (kCacheSlots - 1);
    }
    while (open) {                                    // rare: a third or later probe
        const uint32_t v = C.slot[s];
        if (v == 0xffu) break;
        if (C.hash[v] == h && cache_name_equal(C, v, b, off, len, w0, w1)) { found = (int)v; break; }
        s = (s + 1) & (kCacheSlots - 1);
    }
    base = C.base_off[found < 0 ? 0 : found];
    clen = C.len[found < 0 ? 0 : found];
    if (found < 0) { base = 0; clen = 0; }
    return found;
}
template <class B>
__device__ __forceinline__ int lookup_contig(const ContigCache &C, const DevGenome &g, const B &b, int off, int len,
                                             uint64_t &base, uint64_t &clen)
{
    if (C.n) return cache_find(C, b, off, len, base, clen);
    const int ci = find_contig(g, b, off, len);
    base = ci >= 0 ? g.contigs[ci].base_off : 0;
    clen = ci >= 0 ? g.contigs[ci].len : 0;
    return ci;
}

// A record the tile could not hold (longer than the look-ahead, or longer than
// fgets' 200000-byte buffer): walked from global memory by one thread, split
// the way fgets(buf, MAX_LINE_LEN+1) splits it (pss-bam.c:761-764), counted
// with shared-memory atomics.  Rare by construction.
template <int MODE>
__device__ __noinline__ void long_record(const TallyArgs *Ap, TallyShared *Sp, uint64_t gstart)
{
    const TallyArgs &A = *Ap;
    TallyShared     &S = *Sp;
    uint64_t p = gstart;
    while (p < A.len && __ldg(A.sam + p) != '\n') p++;
    uint64_t total = p - gstart + (p < A.len ? 1u : 0u);
    uint64_t c0 = gstart;
    while (total > 0) {
        const int L = total > (uint64_t)kMaxLine ? kMaxLine : (int)total;
        const GlobalAt at{ A.sam + c0 };
        RecView r;
        int code = scan11(at, L, r);
        if (code == kCounted) {
            uint64_t  cb, cl;
            const int ci = lookup_contig(S.cc, A.g, at, r.rname_off, r.rname_len, cb, cl);
            int code_fk = kFiltered;
            if (MODE != kModePss) {
                FkHits h;
                const TallyCfg &F = (MODE == kModeBoth) ? A.cfg_fk : A.cfg;
                code_fk = fk_record(at, r, true, ci, cb, cl, A.g, F, h);
                if (h.add5) atomicAdd(A.fk_hist + h.idx5, 1ull);
                if (h.add3) atomicAdd(A.fk_hist + (1ull << (2 * F.K)) + h.idx3, 1ull);
                if (MODE == kModeBoth) atomicAdd(&S.stats_fk[stat_slot(code_fk)], 1u);
            }
            if (MODE != kModeFragkon) {
                PssStreams st;
                code = pss_record(at, r, true, ci, cb, cl, A.g, A.cfg, st);
                if (code == kCounted) {
                    const int rows = A.cfg.R + 2;
                    for (int j = 0; j < rows; j++) {
                        if (!((st.a_bad >> (2 * j)) & 1u))
                            atomicAdd(&S.table[j * 16 + (int)(((st.a_read >> (2 * j)) & 3u) * 4 + ((st.a_ref >> (2 * j)) & 3u))], 1u);
                        if (!((st.b_bad >> (2 * j)) & 1u))
                            atomicAdd(&S.table[512 + j * 16 + (int)(((st.b_read >> (2 * j)) & 3u) * 4 + ((st.b_ref >> (2 * j)) & 3u))], 1u);
                    }
                }
            } else {
                code = code_fk;
            }
        } else if (MODE == kModeBoth) {
            atomicAdd(&S.stats_fk[stat_slot(code)], 1u);      // a line that does not parse fails for both programs
        }
        atomicAdd(&S.stats[kStLines], 1u);
        if (MODE == kModeBoth) atomicAdd(&S.stats_fk[kStLines], 1u);
        atomicAdd(&S.stats[stat_slot(code)], 1u);
        log_outcome(A, c0, code);
        c0 += (uint64_t)L;
        total -= (uint64_t)L;
    }
}

// ballot of "(word & mask) != 0" over the full warp: one LOP3-with-predicate + one VOTE
__device__ __forceinline__ uint32_t ballot_bits(uint32_t word, uint32_t mask)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
                 "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t}" : "=r"(r) : "r"(word), "r"(mask));
    return r;
}

// The two count tables of one warp-load of records, without atomics: lane l
// owns cell (l & 15) of table (l >> 4); five ballots per table and row turn
// "which lanes hit my cell" into a popcount.  acc[] packs two rows per
// register (16-bit partial sums).  All 32 lanes must be converged.
template <int NACC>
__device__ __forceinline__ void tally_rows(const PssStreams &st, uint32_t (&acc)[NACC], int rows, uint32_t lane)
{
    const bool     tb = lane >= 16;
    const uint32_t cell = lane & 15u;
    const uint32_t x0 = (cell & 1u) ? 0u : ~0u, x1 = (cell & 2u) ? 0u : ~0u;
    const uint32_t x2 = (cell & 4u) ? 0u : ~0u, x3 = (cell & 8u) ? 0u : ~0u;
    // "row adds nothing" flags become "row counts" flags so that every ballot is a != 0 test
    const uint32_t ar[2] = { (uint32_t)st.a_ref, (uint32_t)(st.a_ref >> 32) }, aq[2] = { (uint32_t)st.a_read, (uint32_t)(st.a_read >> 32) };
    const uint32_t ag[2] = { ~(uint32_t)st.a_bad, ~(uint32_t)(st.a_bad >> 32) };
    const uint32_t br[2] = { (uint32_t)st.b_ref, (uint32_t)(st.b_ref >> 32) }, bq[2] = { (uint32_t)st.b_read, (uint32_t)(st.b_read >> 32) };
    const uint32_t bg[2] = { ~(uint32_t)st.b_bad, ~(uint32_t)(st.b_bad >> 32) };
#pragma unroll
    for (int j = 0; j < 2 * NACC; j++) {
        if (j >= rows) break;
        const int      h = j >> 4;
        const uint32_t m0 = 1u << (2 * (j & 15)), m1 = 2u << (2 * (j & 15));
        const uint32_t a0 = ballot_bits(ar[h], m0), a1 = ballot_bits(ar[h], m1);
        const uint32_t a2 = ballot_bits(aq[h], m0), a3 = ballot_bits(aq[h], m1);
        const uint32_t av = ballot_bits(ag[h], m0);
        const uint32_t b0 = ballot_bits(br[h], m0), b1 = ballot_bits(br[h], m1);
        const uint32_t b2 = ballot_bits(bq[h], m0), b3 = ballot_bits(bq[h], m1);
        const uint32_t bv = ballot_bits(bg[h], m0);
        const uint32_t m = (tb ? bv : av) & ((tb ? b0 : a0) ^ x0) & ((tb ? b1 : a1) ^ x1)
                         & ((tb ? b2 : a2) ^ x2) & ((tb ? b3 : a3) ^ x3);
        acc[j >> 1] += (uint32_t)__popc(m) << (16 * (j & 1));
    }
}
template <int NACC>
__device__ __forceinline__ void flush_acc(uint32_t (&acc)[NACC], int rows, uint32_t lane, uint32_t *table)
{
    const uint32_t tb = lane >> 4, cell = lane & 15u;
#pragma unroll
    for (int j = 0; j < 2 * NACC; j++) {
        if (j >= rows) break;
        const uint32_t v = (acc[j >> 1] >> (16 * (j & 1))) & 0xffffu;
        if (v) atomicAdd(&table[tb * 512 + j * 16 + cell], v);
    }
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0;
}

// ---------------------------------------------------------------------------
// building blocks shared by the two tally kernels
// ---------------------------------------------------------------------------
// zero the CTA tables and mirror the contig table into shared memory (24 human
// chromosomes take < 1 KB).  Called by every thread of the CTA.
__device__ __forceinline__ void cta_prologue(TallyShared &T, const TallyArgs &A, uint32_t tid, uint32_t nthreads)
{
    for (uint32_t i = tid; i < 2 * 32 * 16; i += nthreads) T.table[i] = 0;
    if (tid < 8) { T.stats[tid] = 0; T.stats_fk[tid] = 0; }
    const bool fits = A.g.n_contigs <= (uint32_t)kCacheContigs && A.names_bytes <= (uint32_t)kCacheNames && A.g.n_contigs > 0;
    if (fits) {
        for (uint32_t i = tid; i < (uint32_t)kCacheSlots; i += nthreads) T.cc.slot[i] = 0xffu;
        if (tid == 0) {                          // word offsets of the padded names
            uint32_t w = 0;
            for (uint32_t i = 0; i < A.g.n_contigs; i++) {
                T.cc.name_w0[i] = (uint16_t)w;
                w += (A.g.contigs[i].name_len + 3u) >> 2;
            }
        }
    }
    __syncthreads();
    if (fits && tid < A.g.n_contigs) {
        const DevContig c = A.g.contigs[tid];
        const GlobalAt  nm{ reinterpret_cast<const uint8_t *>(A.g.names) };
        const int       nw = (int)((c.name_len + 3u) >> 2);
        for (int k = 0; k < nw; k++) T.cc.names32[T.cc.name_w0[tid] + k] = field_word(nm, (int)c.name_off, (int)c.name_len, k);
        T.cc.hash[tid] = name_hash_words(nm, (int)c.name_off, (int)c.name_len);
        T.cc.name_len[tid] = (uint16_t)c.name_len;
        T.cc.base_off[tid] = c.base_off; T.cc.len[tid] = c.len;
    }
    __syncthreads();
    if (tid == 0) {
        if (fits)
            for (uint32_t i = 0; i < A.g.n_contigs; i++) {
                uint32_t s = T.cc.hash[i] & (kCacheSlots - 1);
                while (T.cc.slot[s] != 0xffu) s = (s + 1) & (kCacheSlots - 1);
                T.cc.slot[s] = (uint8_t)i;
            }
        T.cc.n = fits ? A.g.n_contigs : 0u;
    }
    __syncthreads();
}
// CTA tables -> global (after a __syncthreads)
template <int MODE>
__device__ __forceinline__ void cta_epilogue(const TallyShared &T, const TallyArgs &A, uint32_t tid, uint32_t nthreads, int rows)
{
    if (MODE != kModeFragkon) {
        for (uint32_t i = tid; i < 2 * 32 * 16; i += nthreads) {
            const uint32_t v = T.table[i];
            const uint32_t tb = i >> 9, row = (i >> 4) & 31u, cell = i & 15u;
            if (v && (int)row < rows) atomicAdd(A.pss_tables + (size_t)tb * rows * 16 + row * 16 + cell, (unsigned long long)v);
        }
    }
    if (tid < kStN && T.stats[tid]) atomicAdd(A.stats + tid, (unsigned long long)T.stats[tid]);
    if (MODE == kModeBoth && tid < kStN && T.stats_fk[tid]) atomicAdd(A.stats_fk + tid, (unsigned long long)T.stats_fk[tid]);
}

// geometry of one tile: smem position p <-> global offset t0 - kPrefix + p
struct TileGeo {
    uint64_t t0;
    int      data_end;      // first smem position past the data
    bool     sees_end;      // the staged span reaches the end of the buffer
};
__device__ __forceinline__ TileGeo tile_geo(const TallyArgs &A, uint64_t tile)
{
    TileGeo g;
    g.t0 = tile * kTileMain;
    const uint64_t avail = A.len - g.t0;                                       // > 0
    g.sees_end = avail <= (uint64_t)(kTileMain + kTileOver);
    g.data_end = kPrefix + (g.sees_end ? (int)avail : kTileMain + kTileOver);
    return g;
}
// one bulk async copy of the tile (plus 16 bytes before it) into `bytes`, completion on `bar`
__device__ __forceinline__ void stage_tile(const TallyArgs &A, const TileGeo &g, uint8_t *bytes, uint64_t *bar)
{
    fence_proxy_async();
    const uint64_t len16 = (A.len + 15) & ~15ull;
    const uint64_t src = g.t0 ? g.t0 - kPrefix : 0;
    uint8_t       *dst = bytes + (g.t0 ? 0 : kPrefix);
    uint64_t       nb = len16 - src;
    const uint64_t cap = (uint64_t)kTileSpan - (g.t0 ? 0 : kPrefix);
    if (nb > cap) nb = cap;
    mbar_expect_tx(bar, (uint32_t)nb);
    bulk_g2s(dst, A.sam + src, (uint32_t)nb, bar);
}
// phase 1: classify bytes, 32 per thread and step -> one le / nl mask word each
__device__ __forceinline__ void classify_tile(const uint8_t *bytes, uint32_t *le, uint32_t *nl, const TileGeo &g,
                                              int first, int stride)
{
    for (int w = first; w < kWords; w += stride) {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(bytes + 32 * w);
        const uint4 v1 = *reinterpret_cast<const uint4 *>(bytes + 32 * w + 16);
        uint32_t zl[8], zn[8];
        classify4(v0.x, zl[0], zn[0]); classify4(v0.y, zl[1], zn[1]);
        classify4(v0.z, zl[2], zn[2]); classify4(v0.w, zl[3], zn[3]);
        classify4(v1.x, zl[4], zn[4]); classify4(v1.y, zl[5], zn[5]);
        classify4(v1.z, zl[6], zn[6]); classify4(v1.w, zl[7], zn[7]);
        uint32_t le32 = (gather8(zl[0], zl[1]) >> 7) | (gather8(zl[2], zl[3]) << 1)
                      | (gather8(zl[4], zl[5]) << 9) | (gather8(zl[6], zl[7]) << 17);
        uint32_t nl32 = (gather8(zn[0], zn[1]) >> 7) | (gather8(zn[2], zn[3]) << 1)
                      | (gather8(zn[4], zn[5]) << 9) | (gather8(zn[6], zn[7]) << 17);
        const int lo = 32 * w;
        if (lo + 32 > g.data_end) {                                            // tail of the data (rare)
            const uint32_t keep = lo >= g.data_end ? 0u : ((1u << (g.data_end - lo)) - 1u);
            le32 &= keep;
            nl32 &= keep;
            if (g.sees_end && g.data_end >= lo && g.data_end < lo + 32) {      // end of buffer terminates the last line
                le32 |= 1u << (g.data_end - lo);
                nl32 |= 1u << (g.data_end - lo);
            }
        }
        le[w] = le32;
        nl[w] = nl32;
    }
}
// newline positions with ordinal in [pass, pass + kRecCap] -> nlpos[]; this thread's words are [w0, w0 + n)
__device__ __forceinline__ void list_newlines(const uint32_t *nl, const uint16_t *wpre, uint16_t *nlpos, int pass, int w0, int n)
{
    for (int k = 0; k < n; k++) {
        if (w0 + k >= kWords) break;
        uint32_t bits = nl[w0 + k];
        int      ord = (int)wpre[w0 + k] - pass;
        while (bits) {
            const int bit = __ffs((int)bits) - 1;
            bits &= bits - 1;
            if (ord >= 0 && ord <= kRecCap) nlpos[ord] = (uint16_t)((w0 + k) * 32 + bit);
            ord++;
        }
    }
}

// One warp-load of records: record i runs from newline i (exclusive) to
// newline i+1 of the pass' list.  Parses, filters, gathers, tallies, counts
// outcomes.  All 32 lanes of the warp must call it together.
template <int MODE, int NACC>
__device__ __forceinline__ void process_batch(const TallyArgs &A, TallyShared &T, const uint8_t *bytes, const uint32_t *le,
                                              const uint16_t *nlpos, int i0, int cnt, int pass, int n_nl, const TileGeo &g,
                                              uint32_t lane, uint32_t (&acc)[NACC], int &acc_iters, int rows)
{
    const uint32_t full = 0xffffffffu;
    const int      i = i0 + (int)lane;
    int            code = 99;                                           // 99 = no record for this lane
    uint64_t       goff = 0;
    int            start = 0, pe = -1;
    if (i < cnt) {
        start = (int)nlpos[i] + 1;
        if (start >= kPrefix && start < kPrefix + kTileMain && start < g.data_end) {
            goff = g.t0 + (uint64_t)(start - kPrefix);
            if (pass + i + 1 < n_nl) { pe = (int)nlpos[i + 1]; code = kNeedSlow; }
            else code = 98;                                             // not whole in this tile
        }
    }
    const SmemAt at{ bytes };
    RecView      r;
    r.flag = 0; r.pos = 0; r.mapq = 0; r.tlen = 0;
    r.rname_off = r.cigar_off = r.seq_off = kPrefix; r.rname_len = r.cigar_len = r.seq_len = 0;
    if (code == kNeedSlow) code = split_fast(at, le, start, pe, r);
    if (code == kNeedSlow) {                                            // rare: glibc sscanf rules
        RecView       rs;
        const SmemRel rel{ bytes + start };
        code = scan11(rel, pe - start, rs);
        r = rs;
        r.rname_off += start; r.cigar_off += start; r.seq_off += start;
    }
    // From here on every lane runs the same straight-line code; lanes whose
    // record did not parse (or that have none) carry valid == false.
    const bool valid = (code == kCounted);
    if (!valid) {                     // a failed parse leaves arbitrary offsets behind: park them on safe bytes
        r.flag = 0; r.pos = 0; r.mapq = 0; r.tlen = 0;
        r.rname_off = r.cigar_off = r.seq_off = kPrefix; r.rname_len = r.cigar_len = r.seq_len = 0;
    }
    uint64_t   cb, cl;
    const int  ci = lookup_contig(T.cc, A.g, at, r.rname_off, valid ? r.rname_len : 0, cb, cl);
    PssStreams st;
    int        code_fk = code;                                  // fragkon's outcome in the fused mode
    if (MODE != kModePss) {
        FkHits          h;
        const TallyCfg &F = (MODE == kModeBoth) ? A.cfg_fk : A.cfg;
        const int       rc = fk_record(at, r, valid, ci, cb, cl, A.g, F, h);
        if (valid) code_fk = rc;
        if (h.add5) atomicAdd(A.fk_hist + h.idx5, 1ull);
        if (h.add3) atomicAdd(A.fk_hist + (1ull << (2 * F.K)) + h.idx3, 1ull);
    }
    if (MODE != kModeFragkon) {
        const int rc = pss_record(at, r, valid, ci, cb, cl, A.g, A.cfg, st);
        if (valid) code = rc;
    } else {
        code = code_fk;
    }
    if (code < 98) log_outcome(A, goff, code);
    __syncwarp();
    if (MODE != kModeFragkon) {
        tally_rows(st, acc, rows, lane);
        if (++acc_iters >= kFlushEvery) { flush_acc(acc, rows, lane, T.table); acc_iters = 0; }
    }
    if (code == 98) { long_record<MODE>(&A, &T, goff); code = 99; code_fk = 99; }
    __syncwarp();
    const uint32_t m_any = __ballot_sync(full, code != 99);
    if (m_any) {
        const uint32_t m0 = __ballot_sync(full, code == kCounted);
        const uint32_t m1 = __ballot_sync(full, code == kNoContig);
        const uint32_t m2 = __ballot_sync(full, code == kFiltered);
        const uint32_t m3 = __ballot_sync(full, code == kParseFail);
        const uint32_t m4 = __ballot_sync(full, code == kUndefined);
        if (lane == 0) {
            atomicAdd(&T.stats[kStLines], (uint32_t)__popc(m_any));
            if (m0) atomicAdd(&T.stats[kStCounted], (uint32_t)__popc(m0));
            if (m1) atomicAdd(&T.stats[kStNoContig], (uint32_t)__popc(m1));
            if (m2) atomicAdd(&T.stats[kStFiltered], (uint32_t)__popc(m2));
            if (m3) atomicAdd(&T.stats[kStParseFail], (uint32_t)__popc(m3));
            if (m4) atomicAdd(&T.stats[kStUndefined], (uint32_t)__popc(m4));
        }
        if (MODE == kModeBoth) {
            const uint32_t f0 = __ballot_sync(full, code_fk == kCounted);
            const uint32_t f1 = __ballot_sync(full, code_fk == kNoContig);
            const uint32_t f2 = __ballot_sync(full, code_fk == kFiltered);
            const uint32_t f3 = __ballot_sync(full, code_fk == kParseFail);
            const uint32_t f4 = __ballot_sync(full, code_fk == kUndefined);
            if (lane == 0) {
                atomicAdd(&T.stats_fk[kStLines], (uint32_t)__popc(m_any));
                if (f0) atomicAdd(&T.stats_fk[kStCounted], (uint32_t)__popc(f0));
                if (f1) atomicAdd(&T.stats_fk[kStNoContig], (uint32_t)__popc(f1));
                if (f2) atomicAdd(&T.stats_fk[kStFiltered], (uint32_t)__popc(f2));
                if (f3) atomicAdd(&T.stats_fk[kStParseFail], (uint32_t)__popc(f3));
                if (f4) atomicAdd(&T.stats_fk[kStUndefined], (uint32_t)__popc(f4));
            }
        }
    }
}

// ---------------------------------------------------------------------------
// tally kernel: one tile at a time per CTA, phases separated by block barriers; four CTAs per SM overlap
// each other's staging, scan and record phases.  (A producer/consumer variant with scan warps and record warps
// handing tiles over through mbarriers was measured and dropped: the record phase is bound by the latency of one
// warp-load of records times the number of such loads resident per SM, which shared memory caps at about 20 either
// way -- profiles/r1_ncu_tally_pipeline_experiment.txt.)
// ---------------------------------------------------------------------------
// NACC = registers of packed partial sums per lane: 9 cover -r <= 16 (the default is 15), 16 cover -r <= 30
template <int MODE, int NACC>
__global__ void __launch_bounds__(kThreads, PSS_TALLY_CTAS_PER_SM) tally_kernel(const __grid_constant__ TallyArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TallySmem &S = *reinterpret_cast<TallySmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t full = 0xffffffffu;

    for (uint32_t i = tid; i < 8; i += kThreads) { S.le[kWords + i] = ~0u; S.nl[kWords + i] = 0u; }
    if (tid == 0) { mbar_init(&S.bar, 1); fence_mbar_init(); }
    cta_prologue(S.sh, A, tid, kThreads);

    const uint64_t n_tiles = (A.len + kTileMain - 1) / kTileMain;
    const int      rows = A.cfg.R + 2;
    uint32_t       phase = 0;
    uint32_t       acc[NACC];
    int            acc_iters = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileGeo g = tile_geo(A, tile);
        if (tid == 0) stage_tile(A, g, S.bytes, &S.bar);
        if (g.t0 == 0 && tid < kPrefix) S.bytes[tid] = '\n';                   // "byte -1" of the stream
        // one warp sleeps on the mbarrier; the others wait at the block barrier without spending issue slots,
        // then take their own (immediately successful) acquire of the completed phase
        if (warp == 0) mbar_wait(&S.bar, phase);
        __syncthreads();
        if (warp != 0) mbar_wait(&S.bar, phase);
        phase ^= 1u;

        classify_tile(S.bytes, S.le, S.nl, g, (int)tid, kThreads);
        __syncthreads();

        // newlines before each mask word (block scan), and -- fused -- the position list of the first pass
        {
            const int w0 = (int)tid * kWordsPerThread;
            uint32_t  nlw[kWordsPerThread], c[kWordsPerThread], sum = 0;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) {
                nlw[k] = (w0 + k < kWords) ? S.nl[w0 + k] : 0u;
                c[k] = (uint32_t)__popc(nlw[k]);
                sum += c[k];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(full, inc, d);
                if ((int)lane >= d) inc += t;
            }
            if (lane == 31) S.warp_sum[warp] = inc;
            __syncthreads();
            uint32_t base = 0;
            if (warp > 0) {                      // sum of the warps below (<= 7 loads, warp uniform)
                const uint32_t ws = lane < warp ? S.warp_sum[lane] : 0u;
                base = __reduce_add_sync(full, ws);
            }
            uint32_t run = base + inc - sum;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) {
                if (w0 + k < kWords) S.wpre[w0 + k] = (uint16_t)run;         // only read again by later passes
                // a 32-byte chunk rarely holds more than one newline: the first one is stored without a branch
                const uint32_t bits = nlw[k];
                if (bits != 0u && run <= (uint32_t)kRecCap) S.nlpos[run] = (uint16_t)((w0 + k) * 32 + __ffs((int)bits) - 1);
                uint32_t rest = bits & (bits - 1u);
                if (__any_sync(full, rest != 0u)) {
                    uint32_t ord = run + 1;
                    while (rest) {
                        if (ord <= (uint32_t)kRecCap) S.nlpos[ord] = (uint16_t)((w0 + k) * 32 + __ffs((int)rest) - 1);
                        rest &= rest - 1u;
                        ord++;
                    }
                }
                run += c[k];
            }
            if (tid == kThreads - 1) S.n_newlines = run;
        }
        __syncthreads();
        const int n_nl = (int)S.n_newlines;

        for (int pass = 0; pass < n_nl; pass += kRecCap) {
            if (pass > 0) {                      // rare: more than kRecCap newlines in one tile
                list_newlines(S.nl, S.wpre, S.nlpos, pass, (int)tid * kWordsPerThread, kWordsPerThread);
                __syncthreads();
            }
            const int cnt = (n_nl - pass) < kRecCap ? (n_nl - pass) : kRecCap;
            for (int i0 = (int)warp * 32; i0 < cnt; i0 += kThreads)
                process_batch<MODE, NACC>(A, S.sh, S.bytes, S.le, S.nlpos, i0, cnt, pass, n_nl, g, lane, acc, acc_iters, rows);
            __syncthreads();             // tile (and nlpos) fully consumed before it is overwritten
        }
        if (n_nl == 0) __syncthreads();
    }

    if (MODE != kModeFragkon) flush_acc(acc, rows, lane, S.sh.table);
    __syncthreads();
    cta_epilogue<MODE>(S.sh, A, tid, kThreads, rows);
}

static_assert(sizeof(TallySmem) + 1024 <= 232448 / PSS_TALLY_CTAS_PER_SM, "the CTAs of the tally kernel must fit one SM");
static_assert(kTileSpan < 65536 && kTileSpan % 16 == 0, "newline positions are kept as u16; bulk copies move 16-byte units");

}  // namespace pssgpu
