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

/* The reference's trie node (kmer.h:9-16).  Declared so that code written against the reference header compiles; this
 * implementation never allocates one. */
typedef struct kmer_tree_node {
    struct kmer_tree_node *Ap, *Cp, *Gp, *Tp;
    unsigned int count;
} ktn;
typedef struct kmer_tree_node *ktnP;

/* First three members: the reference's `Kmers` (kmer.h:18-23), same names, types and offsets.  `ka` is always NULL
 * here -- the counts live in the flat table that follows, so code that walks ka[]/the tries itself (nothing in the
 * reference does outside kmer.c) must go through kmer2count() instead; see INTEGRATION.md "API deviations". */
typedef struct kmers {
    size_t        k;              /* k-mer length */
    size_t        k_ar_size;      /* always K_AR_SIZE, as in the reference */
    ktnP         *ka;             /* reference: 4^8 trie roots; here: NULL */
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
