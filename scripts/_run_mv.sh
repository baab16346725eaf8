#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_pytest.txt 2>&1; tail -3 gpurun_out/r3j_pytest.txt
( for rep in 1 2; do for e in "X=1" "VR_POOL_SLOTS=33554432" "VR_POOL_SLOTS=50331648"; do
  echo "C4 1e9 [$e]: $(env $e python scripts/profile_step.py 1e9 both 2>&1 | tail -1 | cut -d' ' -f6-)"
done; done
for e in "X=1" "VR_POOL_SLOTS=33554432"; do echo "C5 [$e]: $(env $e python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"; done
echo "C4 125e6 [X=1]: $(python scripts/profile_step.py 125e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
echo "C4 125e6 [VR_POOL_SLOTS=33554432]: $(VR_POOL_SLOTS=33554432 python scripts/profile_step.py 125e6 both 2>&1 | tail -1 | cut -d' ' -f6-)" ) > gpurun_out/r3j_pool.txt 2>&1
cat gpurun_out/r3j_pool.txt
