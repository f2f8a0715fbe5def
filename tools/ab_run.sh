#!/bin/bash
# On the GPU box: stage times (tools/stage_times.py) of every library under build/variants/, then of the in-tree one.
cd "$(dirname "$0")/.."
for so in build/variants/*.so; do
  [ -e "$so" ] || continue
  echo "$(basename $so .so): $(SWM_LIB_PATH=$PWD/$so python tools/stage_times.py 2>&1 | tail -1)"
done
echo "in-tree: $(python tools/stage_times.py 2>&1 | tail -1)"
