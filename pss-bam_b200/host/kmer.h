/* Drop-in name of the reference's header; the declarations live in pss_kmer.h. */
#include "pss_kmer.h"
