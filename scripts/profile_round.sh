#!/bin/bash
# One gpurun call that produces the round's measured artefacts under gpurun_out/:
#   bench_<tag>.json            the default bench line (no profiler)
#   launches_<tag>.csv          ncu launch list (gpu__time_duration) of a short bench run
#   prof_<tag>_{trav,shade}.ncu-rep   ncu --set full of one steady-state launch each
tag=${1:-r1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo bench failed; tail -5 gpurun_out/bench_$tag.err; exit 1; }
cat gpurun_out/bench_$tag.json
SHORT="python bench.py --steps 1 --warmup 3 --rays 3e7 --warmup-rays 1e6 --no-cpu-baseline"
$SHORT > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $SHORT > gpurun_out/ncu_launches_$tag.log 2>&1
STEP="python scripts/profile_step.py 64e6 neutral"
$STEP > gpurun_out/plain_step_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 6 -c 1 -f -o gpurun_out/prof_${tag}_trav $STEP > gpurun_out/ncu_trav_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 6 -c 1 -f -o gpurun_out/prof_${tag}_shade $STEP > gpurun_out/ncu_shade_$tag.log 2>&1
tail -2 gpurun_out/plain_step_$tag.log
ls -la gpurun_out | tail -8
