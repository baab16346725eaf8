#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.txt 2>&1; tail -4 gpurun_out/r3a_pytest.txt
( for e in "VR_ENTRY_REFINE=0" "X=1" "VR_ENTRY_REFINE=1" "VR_ENTRY_REFINE=2" "VR_ENTRY_REFINE=8"; do
    echo "C4 both 256e6 [$e]: $(env $e python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
  done
  for e in "VR_ENTRY_REFINE=0" "X=1"; do
    echo "phases lanes1 [$e]: $(env $e VR_LANES=1 VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)"
    echo "C5 4e8 [$e]: $(env $e python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"
    echo "counts [$e]: $(env $e VR_COUNT_WORK=1 python scripts/work_counts.py 2>&1 | tail -2 | tr '\n' ' ')"
  done ) > gpurun_out/r3a_timing.txt 2>&1
cat gpurun_out/r3a_timing.txt
