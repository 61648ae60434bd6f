/* pss_kmer.c -- flat 4^k counter table behind the kmer.h API. */
#include "pss_kmer.h"

KSP init_KSP(int k)
{
    KSP ks;
    if (k < 1 || k > PSS_KMER_MAX_K) return NULL;
    ks = (KSP)malloc(sizeof *ks);
    ks->k = (size_t)k;
    ks->k_ar_size = K_AR_SIZE;
    ks->ka = NULL;
    ks->counts = (unsigned int *)calloc((size_t)1 << (2 * k), sizeof(unsigned int));
    if (!ks->counts) { free(ks); return NULL; }
    return ks;
}

int kmer2inx(const char *kmer, const size_t kmer_len, size_t *inx)
{
    size_t v = 0, i;
    for (i = 0; i < kmer_len; i++) {
        size_t code;
        switch (toupper((unsigned char)kmer[i])) {
        case 'A': code = 0; break;
        case 'C': code = 1; break;
        case 'G': code = 2; break;
        case 'T': code = 3; break;
        default:  return 0;
        }
        v = (v << 2) | code;
    }
    *inx = v;
    return 1;
}

int add_to_ksp(const char *kmer, KSP ks)
{
    size_t inx;
    if (!kmer2inx(kmer, ks->k, &inx)) return -1;
    if (ks->counts[inx] < UINT_MAX) ks->counts[inx]++;
    return 0;
}

unsigned int kmer2count(const char *kmer, const KSP ks)
{
    size_t inx;
    return kmer2inx(kmer, ks->k, &inx) ? ks->counts[inx] : 0u;
}

int ksp_adopt_counts(KSP ks, const uint64_t *counts)
{
    size_t i, n = (size_t)1 << (2 * ks->k);
    for (i = 0; i < n; i++) ks->counts[i] = counts[i] > UINT_MAX ? UINT_MAX : (unsigned int)counts[i];
    return 0;
}

int destroy_KSP(KSP ks)
{
    if (ks) { free(ks->counts); free(ks); }
    return 0;
}
