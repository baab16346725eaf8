#!/bin/bash
# steady-state ncu captures (16.7M-slot launches) of the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEP="python scripts/profile_step.py 256e6 neutral"
$STEP > gpurun_out/r2m_plain.log 2>&1 || { tail -5 gpurun_out/r2m_plain.log; exit 1; }
tail -2 gpurun_out/r2m_plain.log
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 12 -c 1 -f -o gpurun_out/prof_r2m_trav $STEP > gpurun_out/r2m_ncu_trav.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_r2m_shade $STEP > gpurun_out/r2m_ncu_shade.log 2>&1
STEP="python scripts/profile_step.py 256e6 ion"
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_r2m_shade_ion $STEP > gpurun_out/r2m_ncu_shade_ion.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 12 -c 1 -f -o gpurun_out/prof_r2m_trav_ion $STEP > gpurun_out/r2m_ncu_trav_ion.log 2>&1
VR_DUMP_LAUNCHES=1 python scripts/profile_step.py 128e6 both 2>&1 | grep -E "iter|tail" | head -150 > gpurun_out/r2m_dump.txt
ls -la gpurun_out | tail -6
