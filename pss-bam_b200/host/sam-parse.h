/* Drop-in name of the reference's header; the declarations live in pss_sam.h. */
#include "pss_sam.h"
