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

template <bool SMEM>
__global__ void __launch_bounds__(256) spectrum_kernel(const uint64_t *__restrict__ groups, uint64_t g_begin, uint64_t g_end,
                                                       int K, unsigned long long *__restrict__ counts)
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
            else      atomicAdd(counts + idx, 1ull);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(counts + i, (unsigned long long)v);
        }
    }
}

// ===========================================================================
// K1-K4  SAM tile scan + per-record tally
// ===========================================================================
constexpr int kTileMain   = 32768;                            // bytes of SAM a CTA owns per tile
constexpr int kTileOver   = 2048;                             // look-ahead so the last owned record is whole
constexpr int kPrefix     = 16;                               // bytes before the tile (is byte -1 a '\n'?)
constexpr int kTileSpan   = kPrefix + kTileMain + kTileOver;  // 34832 staged bytes
constexpr int kThreads    = 256;
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
    uint16_t name_off[kCacheContigs], name_len[kCacheContigs];
    uint64_t base_off[kCacheContigs], len[kCacheContigs];
    uint8_t  slot[kCacheSlots];                               // open addressing, 0xff = empty
    char     names[kCacheNames];
};

struct TallySmem {
    alignas(128) uint8_t bytes[kWords * 32];                  // staged SAM text
    uint32_t le[kWords + 8];                                  // bit i of word w: byte 32w+i is <= 0x20
    uint32_t nl[kWords + 8];                                  //                  byte 32w+i is '\n'
    uint16_t wpre[kWords + 8];                                // newlines before word w
    uint16_t nlpos[kRecCap + 4];                              // newline positions of this pass
    uint32_t table[2 * 32 * 16];                              // CTA count tables [fwd|rev][row][cell]
    uint32_t stats[8];
    uint32_t warp_sum[kWarps];
    uint32_t n_newlines;
    ContigCache cc;
    alignas(8) uint64_t bar;
};

struct TallyArgs {
    const uint8_t *sam;          // device, 16-byte aligned
    uint64_t       len;
    uint64_t       stream_off;   // offset of sam[0] within everything fed (debug log only)
    DevGenome      g;
    TallyCfg       cfg;
    uint32_t       names_bytes;  // total bytes of contig names
    unsigned long long *pss_tables;   // 2*(R+2)*16 u64: fwd then rev
    unsigned long long *fk_hist;      // 2*4^K u64: 5' then 3'
    unsigned long long *stats;        // kStN
    uint64_t           *dbg_off;      // debug log (may be null)
    int8_t             *dbg_code;
    unsigned long long *dbg_n;
    uint64_t            dbg_cap;
};

struct SmemAt {                  // absolute positions in the staged tile (4-byte aligned base)
    const uint8_t *p;
    __device__ __forceinline__ uint32_t operator()(int i) const { return p[i]; }
    __device__ __forceinline__ uint32_t word(int i) const { return *reinterpret_cast<const uint32_t *>(p + i); }
};
struct SmemRel {                 // positions relative to a record start (any alignment)
    const uint8_t *p;
    __device__ __forceinline__ uint32_t operator()(int i) const { return p[i]; }
    __device__ __forceinline__ uint32_t word(int i) const
    {
        return (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
    }
};
struct GlobalAt {
    const uint8_t *p;
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
    const uint32_t y = w ^ 0x0a0a0a0au;
    znl = ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u;
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

// find_seq (fasta-genome-io.c:202-213) against the shared-memory copy of the contig table
template <class B>
__device__ __forceinline__ int cache_find(const ContigCache &C, const B &b, int off, int len, uint64_t &base, uint64_t &clen)
{
    uint32_t h = kNameHashSeed;
    for (int i = 0; i < len; i++) h = name_hash_step(h, (uint8_t)b(off + i));
    uint32_t s = h & (kCacheSlots - 1);
    int      found = -1;
    for (;;) {
        const uint32_t v = C.slot[s];
        if (v == 0xffu) break;
        if (C.hash[v] == h && (int)C.name_len[v] == len) {
            const char *nm = C.names + C.name_off[v];
            int i = 0;
            while (i < len && (uint8_t)nm[i] == (uint8_t)b(off + i)) i++;
            if (i == len) { found = (int)v; break; }
        }
        s = (s + 1) & (kCacheSlots - 1);
    }
    base = found >= 0 ? C.base_off[found] : 0;
    clen = found >= 0 ? C.len[found] : 0;
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
__device__ __noinline__ void long_record(const TallyArgs *Ap, TallySmem *Sp, uint64_t gstart)
{
    const TallyArgs &A = *Ap;
    TallySmem       &S = *Sp;
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
            if (MODE == kModePss) {
                PssStreams st;
                code = pss_record(at, r, ci, cb, cl, A.g, A.cfg, st);
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
                FkHits h;
                code = fk_record(at, r, ci, cb, cl, A.g, A.cfg, h);
                if (h.add5) atomicAdd(A.fk_hist + h.idx5, 1ull);
                if (h.add3) atomicAdd(A.fk_hist + (1ull << (2 * A.cfg.K)) + h.idx3, 1ull);
            }
        }
        atomicAdd(&S.stats[kStLines], 1u);
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
__device__ __forceinline__ void tally_rows(const PssStreams &st, uint32_t (&acc)[16], int rows, uint32_t lane)
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
    for (int j = 0; j < 32; j++) {
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
__device__ __forceinline__ void flush_acc(uint32_t (&acc)[16], int rows, uint32_t lane, uint32_t *table)
{
    const uint32_t tb = lane >> 4, cell = lane & 15u;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if (j >= rows) break;
        const uint32_t v = (acc[j >> 1] >> (16 * (j & 1))) & 0xffffu;
        if (v) atomicAdd(&table[tb * 512 + j * 16 + cell], v);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) tally_kernel(const __grid_constant__ TallyArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TallySmem &S = *reinterpret_cast<TallySmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t full = 0xffffffffu;

    for (uint32_t i = tid; i < 2 * 32 * 16; i += kThreads) S.table[i] = 0;
    if (tid < 8) S.stats[tid] = 0;
    for (uint32_t i = tid; i < 8; i += kThreads) { S.le[kWords + i] = ~0u; S.nl[kWords + i] = 0u; }
    if (tid == 0) { mbar_init(&S.bar, 1); fence_mbar_init(); }
    // contig table -> shared memory (24 human chromosomes take < 1 KB)
    {
        const bool fits = A.g.n_contigs <= (uint32_t)kCacheContigs && A.names_bytes <= (uint32_t)kCacheNames && A.g.n_contigs > 0;
        if (fits) {
            for (uint32_t i = tid; i < (uint32_t)kCacheSlots; i += kThreads) S.cc.slot[i] = 0xffu;
            for (uint32_t i = tid; i < A.names_bytes; i += kThreads) S.cc.names[i] = A.g.names[i];
            if (tid < A.g.n_contigs) {
                const DevContig c = A.g.contigs[tid];
                uint32_t h = kNameHashSeed;
                for (uint32_t k = 0; k < c.name_len; k++) h = name_hash_step(h, (uint8_t)A.g.names[c.name_off + k]);
                S.cc.hash[tid] = h;
                S.cc.name_off[tid] = (uint16_t)c.name_off; S.cc.name_len[tid] = (uint16_t)c.name_len;
                S.cc.base_off[tid] = c.base_off; S.cc.len[tid] = c.len;
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (fits)
                for (uint32_t i = 0; i < A.g.n_contigs; i++) {
                    uint32_t s = S.cc.hash[i] & (kCacheSlots - 1);
                    while (S.cc.slot[s] != 0xffu) s = (s + 1) & (kCacheSlots - 1);
                    S.cc.slot[s] = (uint8_t)i;
                }
            S.cc.n = fits ? A.g.n_contigs : 0u;
        }
    }
    __syncthreads();

    const uint64_t n_tiles = (A.len + kTileMain - 1) / kTileMain;
    const uint64_t len16 = (A.len + 15) & ~15ull;
    const int      rows = A.cfg.R + 2;
    uint32_t       phase = 0;
    uint32_t       acc[16];
    int            acc_iters = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t t0 = tile * kTileMain;
        // smem position p <-> global offset t0 - kPrefix + p
        const uint64_t avail = A.len - t0;                                     // > 0
        const bool     sees_end = avail <= (uint64_t)(kTileMain + kTileOver);
        const int      data_end = kPrefix + (sees_end ? (int)avail : kTileMain + kTileOver);

        // ---- stage the tile: one bulk async copy, completion on the mbarrier
        if (tid == 0) {
            fence_proxy_async();
            const uint64_t src = t0 ? t0 - kPrefix : 0;
            uint8_t       *dst = S.bytes + (t0 ? 0 : kPrefix);
            uint64_t       nb = len16 - src;
            const uint64_t cap = (uint64_t)kTileSpan - (t0 ? 0 : kPrefix);
            if (nb > cap) nb = cap;
            mbar_expect_tx(&S.bar, (uint32_t)nb);
            bulk_g2s(dst, A.sam + src, (uint32_t)nb, &S.bar);
        }
        if (t0 == 0 && tid < kPrefix) S.bytes[tid] = '\n';                     // "byte -1" of the stream
        mbar_wait(&S.bar, phase);
        phase ^= 1u;
        if (t0 == 0) __syncthreads();

        // ---- phase 1: classify bytes, 32 per thread and step -> one le / nl mask word each
        for (int w = (int)tid; w < kWords; w += kThreads) {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(S.bytes + 32 * w);
            const uint4 v1 = *reinterpret_cast<const uint4 *>(S.bytes + 32 * w + 16);
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
            if (lo + 32 > data_end) {                                          // tail of the data (rare)
                const uint32_t keep = lo >= data_end ? 0u : ((1u << (data_end - lo)) - 1u);
                le32 &= keep;
                nl32 &= keep;
                if (sees_end && data_end >= lo && data_end < lo + 32) {        // end of buffer terminates the last line
                    le32 |= 1u << (data_end - lo);
                    nl32 |= 1u << (data_end - lo);
                }
            }
            S.le[w] = le32;
            S.nl[w] = nl32;
        }
        __syncthreads();

        // ---- phase 1b: newlines before each mask word (block scan)
        {
            const int w0 = (int)tid * kWordsPerThread;
            uint32_t  c[kWordsPerThread], sum = 0;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) {
                c[k] = (w0 + k < kWords) ? (uint32_t)__popc(S.nl[w0 + k]) : 0u;
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
                if (w0 + k < kWords) S.wpre[w0 + k] = (uint16_t)run;
                run += c[k];
            }
            if (tid == kThreads - 1) S.n_newlines = run;
        }
        __syncthreads();
        const int n_nl = (int)S.n_newlines;

        // ---- phase 2: records.  Record i runs from newline i (exclusive) to newline i+1.
        for (int pass = 0; pass < n_nl; pass += kRecCap) {
            {
                const int w0 = (int)tid * kWordsPerThread;
#pragma unroll
                for (int k = 0; k < kWordsPerThread; k++) {
                    if (w0 + k >= kWords) break;
                    uint32_t bits = S.nl[w0 + k];
                    int      ord = (int)S.wpre[w0 + k] - pass;
                    while (bits) {
                        const int bit = __ffs((int)bits) - 1;
                        bits &= bits - 1;
                        if (ord >= 0 && ord <= kRecCap) S.nlpos[ord] = (uint16_t)((w0 + k) * 32 + bit);
                        ord++;
                    }
                }
            }
            __syncthreads();
            const int cnt = (n_nl - pass) < kRecCap ? (n_nl - pass) : kRecCap;
            for (int i0 = (int)warp * 32; i0 < cnt; i0 += kThreads) {
                const int i = i0 + (int)lane;
                int        code = 99;                                   // 99 = no record for this lane
                uint64_t   goff = 0;
                int        start = 0, pe = -1;
                if (i < cnt) {
                    start = (int)S.nlpos[i] + 1;
                    if (start >= kPrefix && start < kPrefix + kTileMain && start < data_end) {
                        goff = t0 + (uint64_t)(start - kPrefix);
                        if (pass + i + 1 < n_nl) { pe = (int)S.nlpos[i + 1]; code = kNeedSlow; }
                        else code = 98;                                 // not whole in this tile
                    }
                }
                const SmemAt at{ S.bytes };
                RecView      r;
                r.flag = 0; r.pos = 0; r.mapq = 0; r.tlen = 0;
                r.rname_off = r.cigar_off = r.seq_off = kPrefix; r.rname_len = r.cigar_len = r.seq_len = 0;
                if (code == kNeedSlow) code = split_fast(at, S.le, start, pe, r);
                if (code == kNeedSlow) {                                // rare: glibc sscanf rules
                    RecView      rs;
                    const SmemRel rel{ S.bytes + start };
                    code = scan11(rel, pe - start, rs);
                    r = rs;
                    r.rname_off += start; r.cigar_off += start; r.seq_off += start;
                }
                PssStreams st;
                st.a_ref = st.a_read = st.b_ref = st.b_read = 0;
                st.a_bad = st.b_bad = kEvenBits;
                if (code == kCounted) {
                    uint64_t  cb, cl;
                    const int ci = lookup_contig(S.cc, A.g, at, r.rname_off, r.rname_len, cb, cl);
                    if (MODE == kModePss) {
                        code = pss_record(at, r, ci, cb, cl, A.g, A.cfg, st);
                    } else {
                        FkHits h;
                        code = fk_record(at, r, ci, cb, cl, A.g, A.cfg, h);
                        if (h.add5) atomicAdd(A.fk_hist + h.idx5, 1ull);
                        if (h.add3) atomicAdd(A.fk_hist + (1ull << (2 * A.cfg.K)) + h.idx3, 1ull);
                    }
                }
                if (code < 98) log_outcome(A, goff, code);
                __syncwarp();
                if (MODE == kModePss) {
                    tally_rows(st, acc, rows, lane);
                    if (++acc_iters >= kFlushEvery) { flush_acc(acc, rows, lane, S.table); acc_iters = 0; }
                }
                if (code == 98) { long_record<MODE>(&A, &S, goff); code = 99; }
                __syncwarp();
                const uint32_t m_any = __ballot_sync(full, code != 99);
                if (m_any) {
                    const uint32_t m0 = __ballot_sync(full, code == kCounted);
                    const uint32_t m1 = __ballot_sync(full, code == kNoContig);
                    const uint32_t m2 = __ballot_sync(full, code == kFiltered);
                    const uint32_t m3 = __ballot_sync(full, code == kParseFail);
                    const uint32_t m4 = __ballot_sync(full, code == kUndefined);
                    if (lane == 0) {
                        atomicAdd(&S.stats[kStLines], (uint32_t)__popc(m_any));
                        if (m0) atomicAdd(&S.stats[kStCounted], (uint32_t)__popc(m0));
                        if (m1) atomicAdd(&S.stats[kStNoContig], (uint32_t)__popc(m1));
                        if (m2) atomicAdd(&S.stats[kStFiltered], (uint32_t)__popc(m2));
                        if (m3) atomicAdd(&S.stats[kStParseFail], (uint32_t)__popc(m3));
                        if (m4) atomicAdd(&S.stats[kStUndefined], (uint32_t)__popc(m4));
                    }
                }
            }
            __syncthreads();             // tile (and nlpos) fully consumed before it is overwritten
        }
        if (n_nl == 0) __syncthreads();
    }

    // ---- warp partial sums -> CTA tables -> global
    if (MODE == kModePss) flush_acc(acc, rows, lane, S.table);
    __syncthreads();
    if (MODE == kModePss) {
        for (uint32_t i = tid; i < 2 * 32 * 16; i += kThreads) {
            const uint32_t v = S.table[i];
            const uint32_t tb = i >> 9, row = (i >> 4) & 31u, cell = i & 15u;
            if (v && (int)row < rows) atomicAdd(A.pss_tables + (size_t)tb * rows * 16 + row * 16 + cell, (unsigned long long)v);
        }
    }
    if (tid < kStN && S.stats[tid]) atomicAdd(A.stats + tid, (unsigned long long)S.stats[tid]);
}

static_assert(sizeof(TallySmem) + 1024 <= 232448 / 4, "four CTAs of the tally kernel must fit one SM");

}  // namespace pssgpu
