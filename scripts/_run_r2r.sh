#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1; tail -4 gpurun_out/r2r_pytest.txt
for r in 125e6 1e9; do
  for l in 1 2 3 4; do
    echo "C4 both rays $r lanes $l: $(VR_LANES=$l python scripts/profile_step.py $r both 2>&1 | tail -1 | cut -d' ' -f6-)"
  done
done > gpurun_out/r2r_lanes.txt 2>&1
for l in 1 2 3; do
  echo "C4 neutral rays 1e8 lanes $l: $(VR_LANES=$l python scripts/profile_step.py 1e8 neutral 2>&1 | tail -1 | cut -d' ' -f6-)"
  echo "C5 rays 4e8 lanes $l: $(VR_LANES=$l python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"
done >> gpurun_out/r2r_lanes.txt 2>&1
cat gpurun_out/r2r_lanes.txt
