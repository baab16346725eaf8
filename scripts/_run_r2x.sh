#!/bin/bash
cd "$(dirname "$0")/.."
for e in "VR_LANES=1 VR_SPREAD_SPLIT=1" "VR_LANES=1 VR_SPREAD_SPLIT=1 VR_NB_SHORTCUT_OFF=1" "VR_LANES=2 VR_SPREAD_SPLIT=1"; do
  echo "C4 both 256e6 [$e]: $(env $e python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
done > gpurun_out/r2x_timing.txt 2>&1
echo "phases lanes1 split: $(VR_SPREAD_SPLIT=1 VR_LANES=1 VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)" >> gpurun_out/r2x_timing.txt
cat gpurun_out/r2x_timing.txt
export VR_LANES=1 VR_SPREAD_SPLIT=1
STEP="python scripts/profile_step.py 256e6 neutral"
ncu --set full --clock-control none --import-source on -k regex:spreadKernel -s 12 -c 1 -f -o gpurun_out/prof_r2x_spread $STEP > gpurun_out/r2x_ncu_spread.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_r2x_shade $STEP > gpurun_out/r2x_ncu_shade.log 2>&1
