#!/bin/bash
# steady-state ncu captures (16.7M-slot launches) of the current build, one lane
cd "$(dirname "$0")/.."
export VR_LANES=1
mkdir -p gpurun_out
STEP="python scripts/profile_step.py 256e6 neutral"
$STEP > gpurun_out/r2t_plain.log 2>&1 || { tail -5 gpurun_out/r2t_plain.log; exit 1; }
tail -2 gpurun_out/r2t_plain.log
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 12 -c 1 -f -o gpurun_out/prof_r2t_trav $STEP > gpurun_out/r2t_ncu_trav.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_r2t_shade $STEP > gpurun_out/r2t_ncu_shade.log 2>&1
STEP="python scripts/profile_step.py 256e6 ion"
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_r2t_shade_ion $STEP > gpurun_out/r2t_ncu_shade_ion.log 2>&1
STEP="python scripts/profile_c5.py 2e8"
$STEP > gpurun_out/r2t_plain_c5.log 2>&1; tail -3 gpurun_out/r2t_plain_c5.log
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 8 -c 1 -f -o gpurun_out/prof_r2t_trav_c5 $STEP > gpurun_out/r2t_ncu_trav_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spreadKernel -s 8 -c 1 -f -o gpurun_out/prof_r2t_spread_c5 $STEP > gpurun_out/r2t_ncu_spread_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 8 -c 1 -f -o gpurun_out/prof_r2t_shade_c5 $STEP > gpurun_out/r2t_ncu_shade_c5.log 2>&1
ls -la gpurun_out | tail -8
