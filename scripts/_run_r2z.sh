#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.txt 2>&1; tail -4 gpurun_out/r2z_pytest.txt
( for r in 256e6 1e9; do echo "C4 both $r: $(python scripts/profile_step.py $r both 2>&1 | tail -1 | cut -d' ' -f6-)"; done
  echo "phases lanes1 256e6: $(VR_LANES=1 VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)"
  echo "C5 4e8: $(python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')" ) > gpurun_out/r2z_timing.txt 2>&1
cat gpurun_out/r2z_timing.txt
