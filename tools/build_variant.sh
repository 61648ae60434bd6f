#!/bin/bash
# developer tool: build a differently tuned libpssgpu (same ABI) into build/variants/<name>.so;  usage: build_variant.sh name -DPSS_...=...
# (select it at run time with PSSGPU_LIB=build/variants/<name>.so)
set -e
name=$1; shift
cd "$(dirname "$0")/../pss-bam_b200/csrc"
mkdir -p ../../build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -shared -Xcompiler -fPIC --threads 3 "$@" \
    -o ../../build/variants/$name.so pssgpu.cu pssgpu_bam.cu pssgpu_group.cu -ldl
