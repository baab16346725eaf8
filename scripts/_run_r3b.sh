#!/bin/bash
cd "$(dirname "$0")/.."
for v in entry; do for e in "VR_ENTRY_REFINE=0" "VR_ENTRY_REFINE=4"; do echo "== $v [$e]"; env $e VR_LIB_PATH=$PWD/variants/$v.so python scripts/exp_homog.py 4e6 2>&1 | tail -5; done; done > gpurun_out/r3b_homog.txt 2>&1
cat gpurun_out/r3b_homog.txt
