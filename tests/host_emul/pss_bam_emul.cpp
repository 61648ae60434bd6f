// Host build of the device BGZF/BAM ingest logic (pss_inflate.h, pss_bamrec.h) for the CPU test-suite.  A TEST of
// that logic: the product executes it only inside CUDA kernels (pssgpu_bam.cu).
#include <cstring>
#include <vector>

#include "../../pss-bam_b200/csrc/pss_bamrec.h"
#include "../../pss-bam_b200/csrc/pss_crc32.h"
#include "../../pss-bam_b200/csrc/pss_inflate.h"

using namespace pssgpu;

extern "C" {

int emul_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len)
{
    static thread_local InflateTables T;
    return inflate_block(in, in_len, out, out_len, T);
}

// CRC-32 of [p, p + n) the way the inflate kernel computes it: the 32 lanes of the warp played one after the other
uint32_t emul_crc32(const uint8_t *p, uint32_t n)
{
    static uint32_t tab[kCrcTableWords];
    static bool     built = false;
    if (!built) { crc32_build_tables(tab); built = true; }
    uint32_t       head;
    uint32_t       s = crc32_head(p, n, tab, &head);
    const uint32_t n_words = (n - head) >> 2;
    if (n_words) {
        uint32_t c = 0;
        for (uint32_t lane = 0; lane < (uint32_t)kCrcLanes; lane++) {
            uint32_t       shift;
            const uint32_t u = crc32_lane_partial(reinterpret_cast<const uint32_t *>(p + head), n_words, lane, s, tab, &shift);
            if (shift) c ^= crc_mulmod(tab[kCrcXk + shift], u);
        }
        s = c;
    }
    return crc32_tail(p + head + 4u * n_words, (n - head) & 3u, s, tab);
}

// header of an inflated BAM stream: returns its length (0: incomplete / not BAM) and fills the dictionary
static uint64_t parse_header(const uint8_t *u, uint64_t len, std::vector<uint32_t> &off, std::vector<uint32_t> &nl)
{
    if (len < 12 || bam_u32(u) != 0x014d4142u) return 0;
    const uint64_t l_text = bam_u32(u + 4);
    if (len < 12 + l_text) return 0;
    const uint32_t n_ref = bam_u32(u + 8 + l_text);
    uint64_t q = 12 + l_text;
    for (uint32_t i = 0; i < n_ref; i++) {
        if (q + 4 > len) return 0;
        const uint64_t l_name = bam_u32(u + q);
        if (q + 8 + l_name > len) return 0;
        off.push_back((uint32_t)(q + 4));
        nl.push_back(l_name ? (uint32_t)(l_name - 1) : 0u);
        q += 8 + l_name;
    }
    return q;
}

// Whole inflated BAM stream -> the text the render kernel produces (records in file order).  rg: NULL = no filter.
// Returns the text length, or -1 (malformed) / -2 (capacity).  u must have 8 bytes of slack behind len.
long emul_bam_render(const uint8_t *u, uint64_t len, const char *rg, uint8_t *text, uint64_t cap, uint64_t *n_records, uint64_t *n_dropped)
{
    std::vector<uint32_t> off, nl;
    const uint64_t h = parse_header(u, len, off, nl);
    if (!h) return -1;
    BamRefs R{ u, off.data(), nl.data(), (int32_t)off.size() };
    const int rg_len = rg ? (int)strlen(rg) : -1;
    uint64_t  x = h, at = 0;
    *n_records = *n_dropped = 0;
    while (x < len) {
        if (x + 4 + kBamFixed > len) return -1;
        const BamCore c = bam_core(u + x);
        if (!bam_wellformed(c) || x + 4 + c.block_size > len) return -1;
        (*n_records)++;
        if (rg_len >= 0 && !bam_has_read_group(u + x, c, rg, rg_len)) (*n_dropped)++;
        else {
            BamCountSink cs;
            bam_render(u + x, c, R, cs);
            if (at + cs.n > cap) return -2;
            BamWriteSink ws{ text + at };
            bam_render(u + x, c, R, ws);
            if ((uint64_t)(ws.p - (text + at)) != cs.n) return -3;
            at += cs.n;
        }
        x += 4ull + c.block_size;
    }
    return (long)at;
}

// The first-record guess of the guess kernel at every multiple of `block` behind the header, against the true chain.
// out[0] = boundaries tested, out[1] = guesses that differ from the truth, out[2] = boundaries with no guess at all.
int emul_bam_guess_stats(const uint8_t *u, uint64_t len, uint32_t block, uint64_t *out)
{
    std::vector<uint32_t> off, nl;
    const uint64_t h = parse_header(u, len, off, nl);
    if (!h) return -1;
    std::vector<uint64_t> starts;
    for (uint64_t x = h; x + 4 <= len;) { starts.push_back(x); x += 4ull + bam_u32(u + x); }
    out[0] = out[1] = out[2] = 0;
    size_t k = 0;
    for (uint64_t u0 = ((h / block) + 1) * block; u0 < len; u0 += block) {
        const uint64_t u1 = u0 + block < len ? u0 + block : len;
        while (k < starts.size() && starts[k] < u0) k++;
        const uint64_t truth = (k < starts.size() && starts[k] < u1) ? starts[k] : ~0ull;
        uint64_t guess = ~0ull;
        for (uint64_t x = u0; x < u1; x++)
            if (bam_guess_at(u, x, len, (int32_t)off.size())) { guess = x; break; }
        out[0]++;
        if (guess != truth) out[1]++;
        if (guess == ~0ull && truth != ~0ull) out[2]++;
    }
    return 0;
}

}  // extern "C"
