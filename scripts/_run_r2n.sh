#!/bin/bash
cd "$(dirname "$0")/.."
RAYS=256e6 WHICH=neutral bash scripts/exp_variants.sh base pfkids pfpush pfleaf1 pfleaf2 pfleaf4 pfpop disk1 disk2 node1 node1disk1 pfall base > gpurun_out/r2n_variants.txt 2>&1
cat gpurun_out/r2n_variants.txt
