#!/bin/bash
# A/B of the pyramid kernel variants on the GPU box: rebuilds libswm_orb.so with -D switches and times the stages.
cd "$(dirname "$0")/.."
for v in "1 1" "0 1" "1 0" "0 0"; do
  set -- $v
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -ldl -DSWM_PYR_TMA=$1 -DSWM_PYR_PACKED=$2 \
    -o swarmmap_b200/libswm_orb.so swarmmap_b200/csrc/extract.cu swarmmap_b200/csrc/match.cu swarmmap_b200/csrc/bow.cu || exit 1
  echo "TMA=$1 PACKED=$2: $(python tools/stage_times.py)"
done
