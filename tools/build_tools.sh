#!/bin/bash
# Builds the stand-alone kernel bring-up tests into tools/_bin (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
mkdir -p _bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DBIGRU_NO_TRAP ${TOOLS_DEFS:-}"
for t in ${@:-tc_scanx_test tc_scanw_test tc_scan_test tc_gemm_test mma_bench}; do
  nvcc $FLAGS -o _bin/$t $t.cu
done
