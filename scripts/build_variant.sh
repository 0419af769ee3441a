#!/bin/bash
# scripts/build_variant.sh NAME "<extra nvcc flags>"  ->  diffdart_b200/lib/variants/libddb_NAME.so
# Rebuilds only the dynamics translation unit (ddb_api.cu) with extra -D flags and links it with the current lcp_api.o:
# for A/B timing of the contact-free kernels on the GPU box (DDB_LIB=... python scripts/kbench.py).
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p $HERE/diffdart_b200/lib/variants /tmp/ddbvar
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -diag-suppress 549 \
  --split-compile 0 $@ -c -o /tmp/ddbvar/ddb_api_$NAME.o $HERE/diffdart_b200/csrc/ddb_api.cu 2> $HERE/diffdart_b200/lib/variants/ptxas_$NAME.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $HERE/diffdart_b200/lib/variants/libddb_$NAME.so /tmp/ddbvar/ddb_api_$NAME.o $HERE/diffdart_b200/lib/lcp_api.o $HERE/diffdart_b200/lib/lemke_api.o
grep -A3 "k_forwardILi5ELb1\|k_backwardILi5ELb1" $HERE/diffdart_b200/lib/variants/ptxas_$NAME.log | grep "spill\|Used" 
