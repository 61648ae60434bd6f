// pss_emul.cpp -- TEST HARNESS, not product code.
//
// Compiles pss-bam_b200/csrc/pss_record.h (the per-record logic the CUDA
// kernels run) for the host, so that the CPU test-suite can fuzz it against
// the oracle without a GPU: sscanf emulation, clean-record splitter, filters,
// packed-genome windows, strand handling, fragkon windows.  The tile scan,
// the ballot tally and the histogram atomics are device-only and are covered
// by the `-m gpu` parity tests.  Nothing under pss-bam_b200/ links this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../pss-bam_b200/csrc/pss_record.h"

using namespace pssgpu;

namespace {

struct HostGenome {
    std::vector<uint64_t>  groups;
    std::vector<DevContig> contigs;
    std::string            names;
    std::vector<uint32_t>  hash;
    std::vector<uint64_t>  exc_pos;
    std::vector<uint8_t>   exc_chr;
    DevGenome view() const
    {
        DevGenome g;
        g.groups = groups.data(); g.n_groups = groups.size();
        g.contigs = contigs.data(); g.n_contigs = (uint32_t)contigs.size();
        g.names = names.data(); g.hash = hash.data(); g.hash_mask = (uint32_t)hash.size() - 1;
        g.exc_pos = exc_pos.data(); g.exc_chr = exc_chr.data(); g.n_exc = (uint32_t)exc_pos.size();
        return g;
    }
};

struct At {
    static constexpr int kLookBack = 0;
    const uint8_t *p;
    int lo() const { return 0; }
    uint32_t operator()(int i) const { return p[i]; }
    uint32_t word(int i) const { return (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24); }
};

void fill_ctx(const char *s, uint32_t *mask, uint32_t *other, char *copy)
{
    static const char named[] = PSSGPU_SYM_CHARS;
    *mask = 0; *other = 0;
    memset(copy, 0, kMaxCtxChars);
    strncpy(copy, s, kMaxCtxChars - 1);
    for (size_t i = 0; s[i]; i++) {
        const char *p = strchr(named, s[i]);
        if (p && *p) *mask |= 1u << (p - named); else *other = 1;
    }
}

}  // namespace

extern "C" {

void *emul_genome_new(const char *const *ids, const char *const *seqs, const uint64_t *lens, uint32_t n)
{
    HostGenome *G = new HostGenome();
    uint64_t cur = kPadBases;
    for (uint32_t i = 0; i < n; i++) {
        DevContig c;
        c.base_off = cur; c.len = lens[i];
        c.name_off = (uint32_t)G->names.size(); c.name_len = (uint32_t)strlen(ids[i]);
        G->names += ids[i];
        G->contigs.push_back(c);
        cur += ((lens[i] + 15) / 16) * 16 + kPadBases;
    }
    G->groups.assign(cur / 16 + 2, ~0ull);
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t *s = (const uint8_t *)seqs[i];
        for (uint64_t b0 = 0; b0 < lens[i]; b0 += 16) {
            int nv = lens[i] - b0 >= 16 ? 16 : (int)(lens[i] - b0);
            uint32_t om = 0, nul = 0;
            uint64_t grp = pack_group([&](int j) { return s[b0 + j]; }, nv, &om, &nul);
            G->groups[(G->contigs[i].base_off + b0) / 16] = grp;
            for (int j = 0; j < 16; j++)
                if (om & (1u << j)) { G->exc_pos.push_back(G->contigs[i].base_off + b0 + j); G->exc_chr.push_back(upper_c(s[b0 + j])); }
        }
    }
    uint32_t hs = 16;
    while (hs < 2 * n + 1) hs <<= 1;
    G->hash.assign(hs, 0);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t h = kNameHashSeed;
        for (uint32_t k = 0; k < G->contigs[i].name_len; k++) h = name_hash_step(h, (uint8_t)G->names[G->contigs[i].name_off + k]);
        uint32_t slot = h & (hs - 1);
        while (G->hash[slot]) slot = (slot + 1) & (hs - 1);
        G->hash[slot] = i + 1;
    }
    return G;
}
void emul_genome_free(void *g) { delete (HostGenome *)g; }

// sscanf emulation alone: returns the status and the converted values
int emul_scan11(const char *line, int L, uint32_t *flag, uint64_t *pos, uint32_t *mapq, int32_t *tlen,
                int32_t *offs /* rname_off,len, cigar_off,len, seq_off,len */)
{
    RecView r;
    memset(&r, 0, sizeof r);
    int code = scan11(At{ (const uint8_t *)line }, L, r);
    *flag = r.flag; *pos = r.pos; *mapq = r.mapq; *tlen = r.tlen;
    offs[0] = r.rname_off; offs[1] = r.rname_len; offs[2] = r.cigar_off; offs[3] = r.cigar_len;
    offs[4] = r.seq_off; offs[5] = r.seq_len;
    return code;
}

// The whole per-record path over a SAM block.  mode 0 = pss-bam, 1 = fragkon.
// force_slow != 0 skips split_fast (every line through scan11).
// Returns the number of lines; `used_fast` counts lines split_fast accepted.
uint64_t emul_tally(void *gp, const char *sam, uint64_t len, int mode,
                    int R, uint64_t min_len, uint64_t max_len, int min_mq, const char *up, const char *down,
                    int merged_only, int K, int force_slow,
                    uint64_t *fwd, uint64_t *rev, uint64_t *fp, uint64_t *tp,
                    int8_t *status, uint64_t status_cap, uint64_t *used_fast)
{
    const HostGenome *G = (const HostGenome *)gp;
    const DevGenome   g = G->view();
    TallyCfg P;
    memset(&P, 0, sizeof P);
    P.mode = mode; P.R = R; P.min_len = min_len; P.max_len = max_len; P.min_mq = (uint32_t)min_mq;
    P.merged_only = merged_only ? 1u : 0u; P.K = K;
    fill_ctx(up ? up : "ACGT", &P.up_mask, &P.up_other, P.up_ctx);
    fill_ctx(down ? down : "ACGT", &P.down_mask, &P.down_other, P.down_ctx);

    // separator mask the way the tile scan builds it (+ virtual terminator at len)
    std::vector<uint32_t> le((len + 32) / 32 + 12, 0);
    for (uint64_t i = 0; i < len; i++)
        if ((uint8_t)sam[i] <= 0x20) le[i >> 5] |= 1u << (i & 31);
    le[len >> 5] |= 1u << (len & 31);
    for (size_t k = (len >> 5) + 1; k < le.size(); k++) le[k] = ~0u;

    const uint8_t *sb = (const uint8_t *)sam;
    uint64_t nlines = 0, fast = 0, start = 0;
    while (start < len) {
        uint64_t pe = start;
        while (pe < len && sb[pe] != '\n') pe++;
        uint64_t total = pe - start + (pe < len ? 1 : 0);
        uint64_t c0 = start;
        while (total > 0) {              // fgets(buf, 200001): pss-bam.c:761-764
            const int L = total > (uint64_t)kMaxLine ? kMaxLine : (int)total;
            const bool whole = (c0 == start && (uint64_t)L == total);
            RecView r;
            memset(&r, 0, sizeof r);
            const At at{ sb + c0 };
            int code = kNeedSlow;
            if (whole && !force_slow && len < 0x7fffff00ull) {
                const At abs0{ sb };
                code = split_fast(abs0, le.data(), (int)start, (int)pe, r);
                if (code != kNeedSlow) {
                    fast++;
                    r.rname_off -= (int)start; r.cigar_off -= (int)start; r.seq_off -= (int)start;
                }
            }
            if (code == kNeedSlow) code = scan11(at, L, r);
            if (code == kCounted) {
                const int ci = find_contig(g, at, r.rname_off, r.rname_len);
                const uint64_t cb = ci >= 0 ? g.contigs[ci].base_off : 0, cl = ci >= 0 ? g.contigs[ci].len : 0;
                if (mode == kModePss && R > kMaxRegion) {
                    code = pss_record_wide(at, r, ci, cb, cl, g, P, [&](int tb, int row, int cell) { (tb ? rev : fwd)[row * 16 + cell]++; });
                } else if (mode == kModePss) {
                    PssStreams st;
                    code = pss_record(at, r, true, ci, cb, cl, g, P, st);
                    if (code == kCounted) {
                        for (int j = 0; j < R + 2; j++) {
                            if (!((st.a_bad >> (2 * j)) & 1u)) fwd[j * 16 + ((st.a_read >> (2 * j)) & 3u) * 4 + ((st.a_ref >> (2 * j)) & 3u)]++;
                            if (!((st.b_bad >> (2 * j)) & 1u)) rev[j * 16 + ((st.b_read >> (2 * j)) & 3u) * 4 + ((st.b_ref >> (2 * j)) & 3u)]++;
                        }
                    }
                } else {
                    FkHits h;
                    code = fk_record(at, r, true, ci, cb, cl, g, P, h);
                    if (h.add5) fp[h.idx5]++;
                    if (h.add3) tp[h.idx3]++;
                }
            }
            if (status && nlines < status_cap) status[nlines] = (int8_t)code;
            nlines++;
            c0 += (uint64_t)L;
            total -= (uint64_t)L;
        }
        start = pe + 1;
    }
    if (used_fast) *used_fast = fast;
    return nlines;
}

// k-mer spectrum over the packed groups exactly as spectrum_kernel walks them
void emul_spectrum(void *gp, int K, uint64_t *counts)
{
    const HostGenome *G = (const HostGenome *)gp;
    const uint32_t kmask = (1u << (2 * K)) - 1u;
    for (size_t gi = 0; gi + 2 < G->groups.size(); gi++) {
        const uint64_t g0 = G->groups[gi], g1 = G->groups[gi + 1];
        const uint64_t codes = (uint64_t)(uint32_t)g0 | ((uint64_t)(uint32_t)g1 << 32);
        const uint64_t cls = (g0 >> 32) | (g1 & 0xffffffff00000000ull);
        const uint64_t bad = (cls | (cls >> 1)) & kEvenBits;
        for (int o = 0; o < 16; o++) {
            if (((bad >> (2 * o)) & (uint64_t)kmask) != 0) continue;
            const uint32_t w = (uint32_t)(codes >> (2 * o)) & kmask;
            counts[rev_fields32(w) >> (32 - 2 * K)]++;
        }
    }
}

}  // extern "C"
