/* Drop-in name of the reference's header; the declarations live in pss_fasta.h. */
#include "pss_fasta.h"
