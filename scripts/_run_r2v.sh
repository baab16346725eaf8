#!/bin/bash
cd "$(dirname "$0")/.."
export VR_LANES=1
RAYS=256e6 WHICH=neutral bash scripts/exp_variants.sh new sc sc_noq sc_off > gpurun_out/r2v_variants.txt 2>&1
cat gpurun_out/r2v_variants.txt
STEP="python scripts/profile_step.py 256e6 neutral"
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_r2v_shade $STEP > gpurun_out/r2v_ncu_shade.log 2>&1
