#!/bin/bash
# builds tuning variants of the library into variants/<name>.so:  name:"-DX=.. -DY=.."
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -shared"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  ( /usr/local/cuda/bin/nvcc $FLAGS $defs viennaray_b200/csrc/vr_api.cu viennaray_b200/csrc/vr_bvh.cu viennaray_b200/csrc/vr_scene.cu viennaray_b200/csrc/vr_trace.cu -o variants/$name.so && echo built $name ) &
done
wait
