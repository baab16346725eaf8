#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2y_pytest.txt 2>&1; tail -4 gpurun_out/r2y_pytest.txt
export VR_LANES=1
( for v in sc3 sc3_noq; do
  for e in "X=1" "VR_SPREAD_SPLIT=1" "VR_NB_SHORTCUT_OFF=1"; do
    echo "$v C4 both 256e6 [$e]: $(env $e VR_LIB_PATH=$PWD/variants/$v.so python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
  done; done
  echo "phases inline: $(VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)"
  echo "phases split: $(VR_SPREAD_SPLIT=1 VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)"
  for e in "X=1" "VR_NB_SHORTCUT_OFF=1"; do
    echo "C5 4e8 [$e]: $(env $e python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"
  done ) > gpurun_out/r2y_timing.txt 2>&1
cat gpurun_out/r2y_timing.txt
