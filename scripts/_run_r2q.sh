#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.txt 2>&1; tail -4 gpurun_out/r2q_pytest.txt
for r in 125e6 256e6 1e9; do
  for l in 1 2; do
    echo "rays $r lanes $l: $(VR_LANES=$l python scripts/profile_step.py $r both 2>&1 | tail -1)"
  done
done > gpurun_out/r2q_lanes.txt 2>&1
cat gpurun_out/r2q_lanes.txt
