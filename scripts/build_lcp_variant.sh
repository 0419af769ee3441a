#!/bin/bash
# scripts/build_lcp_variant.sh NAME "<extra nvcc flags>"  ->  diffdart_b200/lib/variants/libddb_NAME.so
# Rebuilds only the contact translation unit (lcp_api.cu) with extra -D flags and links it with the current ddb_api.o:
# for A/B timing of the contact kernels on the GPU box (DDB_LIB=... python scripts/bench_contact.py).
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p $HERE/diffdart_b200/lib/variants /tmp/ddbvar
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -diag-suppress 549,177 \
  -fmad=false $@ -c -o /tmp/ddbvar/lcp_api_$NAME.o $HERE/diffdart_b200/csrc/lcp_api.cu 2> $HERE/diffdart_b200/lib/variants/ptxas_lcp_$NAME.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $HERE/diffdart_b200/lib/variants/libddb_$NAME.so $HERE/diffdart_b200/lib/ddb_api.o /tmp/ddbvar/lcp_api_$NAME.o $HERE/diffdart_b200/lib/lemke_api.o
grep -A2 "k_contactE\|k_contact_backwardILi8E" $HERE/diffdart_b200/lib/variants/ptxas_lcp_$NAME.log | grep "spill\|Used"
