#!/bin/bash
cd "$(dirname "$0")/.."
RAYS=256e6 WHICH=both bash scripts/exp_variants.sh new new_b5 new_b3 new_b6 > gpurun_out/r2p_variants.txt 2>&1
cat gpurun_out/r2p_variants.txt
( echo "== 1 ctx"; python scripts/exp_corun.py 128e6 1 2>&1 | tail -2
  echo "== 2 ctx no caps"; python scripts/exp_corun.py 128e6 2 2>&1 | tail -2
  echo "== 2 ctx cap5 pad100k"; VR_TRAV_CAP=5 VR_SHADE_SMEM_PAD=102400 python scripts/exp_corun.py 128e6 2 2>&1 | tail -2
  echo "== 2 ctx cap7 pad100k"; VR_TRAV_CAP=7 VR_SHADE_SMEM_PAD=102400 python scripts/exp_corun.py 128e6 2 2>&1 | tail -2
  echo "== 2 ctx cap6 pad70k"; VR_TRAV_CAP=6 VR_SHADE_SMEM_PAD=71680 python scripts/exp_corun.py 128e6 2 2>&1 | tail -2
  echo "== 2 ctx cap5 nopad"; VR_TRAV_CAP=5 python scripts/exp_corun.py 128e6 2 2>&1 | tail -2
  echo "== 1 ctx cap5 pad100k"; VR_TRAV_CAP=5 VR_SHADE_SMEM_PAD=102400 python scripts/exp_corun.py 128e6 1 2>&1 | tail -2
  echo "== 3 ctx cap3 pad100k"; VR_TRAV_CAP=3 VR_SHADE_SMEM_PAD=102400 VR_POOL_SLOTS=8388608 python scripts/exp_corun.py 96e6 3 2>&1 | tail -2
) > gpurun_out/r2p_corun.txt 2>&1
cat gpurun_out/r2p_corun.txt
