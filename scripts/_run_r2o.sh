#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.txt 2>&1; tail -5 gpurun_out/r2o_pytest.txt
RAYS=256e6 WHICH=both bash scripts/exp_variants.sh base new new:VR_BOUNDARY_GENERIC=1 new_csr new_csr:VR_BOUNDARY_GENERIC=1 base > gpurun_out/r2o_variants.txt 2>&1
cat gpurun_out/r2o_variants.txt
