#!/bin/bash
cd "$(dirname "$0")/.."
( for rep in 1 2; do for v in k0 k_phil k_bnd; do
  echo "$v: $(VR_LIB_PATH=$PWD/variants/$v.so python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
done
for e in "VR_TAIL_RAYS=131072" "VR_TAIL_RAYS=524288" "VR_TAIL_RAYS=1048576"; do
  echo "k0 [$e] 256e6: $(env $e VR_LIB_PATH=$PWD/variants/k0.so python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
  echo "k0 [$e] 125e6: $(env $e VR_LIB_PATH=$PWD/variants/k0.so python scripts/profile_step.py 125e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
done; done
echo "k0 125e6: $(VR_LIB_PATH=$PWD/variants/k0.so python scripts/profile_step.py 125e6 both 2>&1 | tail -1 | cut -d' ' -f6-)" ) > gpurun_out/r3k_micro.txt 2>&1
cat gpurun_out/r3k_micro.txt
