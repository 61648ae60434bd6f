#!/bin/sh
# Test shim for `samtools view` (samtools is absent from this image).  The
# reference programs popen("samtools view [-r RG] FILE") and read SAM text;
# our "BAM" test inputs are plain SAM text files, so:
#   samtools view FILE        -> FILE without '@' header lines
#   samtools view -r RG FILE  -> same, restricted to records tagged RG:Z:RG
# Not a product component.
[ "$1" = "view" ] || { echo "samtools shim: only 'view' is supported" >&2; exit 1; }
shift
RG=""
if [ "$1" = "-r" ]; then RG="$2"; shift 2; fi
if [ -z "$RG" ]; then
  exec grep -v '^@' "$1"
else
  grep -v '^@' "$1" | grep -P "\tRG:Z:${RG}(\t|\$)"
fi
