#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do for v in fin2 fin2_leaf10; do
  echo "$v C5: $(VR_LIB_PATH=$PWD/variants/$v.so python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"
  echo "$v C4 1e9: $(VR_LIB_PATH=$PWD/variants/$v.so python scripts/profile_step.py 1e9 both 2>&1 | tail -1 | cut -d' ' -f6-)"
done; done > gpurun_out/r3g_leaf.txt 2>&1
cat gpurun_out/r3g_leaf.txt
