#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.txt 2>&1; tail -6 gpurun_out/r2w_pytest.txt
for e in "VR_LANES=1" "VR_LANES=1 VR_NB_SHORTCUT_OFF=1" "VR_LANES=2"; do
  echo "C4 both 256e6 [$e]: $(env $e python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
done > gpurun_out/r2w_timing.txt 2>&1
echo "phases lanes1: $(VR_LANES=1 VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)" >> gpurun_out/r2w_timing.txt
for e in "VR_LANES=2" "VR_LANES=2 VR_NB_SHORTCUT_OFF=1"; do
  echo "C5 4e8 [$e]: $(env $e python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"
done >> gpurun_out/r2w_timing.txt 2>&1
cat gpurun_out/r2w_timing.txt
