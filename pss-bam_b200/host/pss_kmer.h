/* pss_kmer.h -- k-mer count tables of the host side.
 *
 * API of the reference's kmer.h:27-33 (`kmer.h` here includes this file).
 * The reference keeps an array over the first 8 bases whose slots root 4-ary
 * pointer tries for the remaining bases (kmer.h:9-24, kmer.c:43-110); this
 * implementation keeps one flat table of 4^k saturating counters indexed by
 * the 2-bit code of the k-mer (A=0 C=1 G=2 T=3, first base most significant,
 * kmer.c:184-214) -- the layout the GPU histograms use, so device results can
 * be adopted without conversion (ksp_adopt_counts).  Observable behaviour is
 * the same: non-ACGT bytes reject a k-mer, counts stop at UINT_MAX.
 */
#ifndef PSS_KMER_H
#define PSS_KMER_H

#include <ctype.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define K_AR_SIZE (8)
#define PSS_KMER_MAX_K (15)       /* 4^15 counters = 4 GiB of unsigned int */

typedef struct kmers {
    size_t        k;              /* k-mer length */
    size_t        k_ar_size;      /* kept for source compatibility (always K_AR_SIZE) */
    unsigned int *counts;         /* 4^k saturating counters */
} Kmers;
typedef struct kmers *KSP;

KSP          init_KSP(int k);                                  /* NULL when k is outside [1, PSS_KMER_MAX_K] */
int          add_to_ksp(const char *kmer, KSP ks);             /* 0 counted, -1 not a valid k-mer */
unsigned int kmer2count(const char *kmer, const KSP ks);       /* 0 for invalid k-mers */
int          kmer2inx(const char *kmer, const size_t kmer_len, size_t *inx);   /* 1 valid, 0 invalid */
int          destroy_KSP(KSP ks);

/* extension: fill the table from 4^k unsaturated 64-bit counts (device output), clamping at UINT_MAX */
int          ksp_adopt_counts(KSP ks, const uint64_t *counts);

#endif
