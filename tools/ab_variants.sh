#!/bin/bash
# Builds kernel variants of libswm_orb.so HERE (nvcc cross-compiles without a GPU) into build/variants/<name>.so:
#   tools/ab_variants.sh name "-DSWITCH=1 ..." [name2 "..."]...
# and tools/ab_run.sh times them on the GPU box (SWM_LIB_PATH selects the library).
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -ldl $2 \
    -o build/variants/$1.so swarmmap_b200/csrc/extract.cu swarmmap_b200/csrc/match.cu swarmmap_b200/csrc/bow.cu || exit 1
  echo "built build/variants/$1.so ($2)"
  shift 2
done
