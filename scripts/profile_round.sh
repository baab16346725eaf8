#!/bin/bash
# One gpurun call that produces a round's measured artefacts under gpurun_out/ (tag = $1):
#   <tag>_pytest.txt, <tag>_smoke.txt     pytest -m gpu, __graft_entry__.smoke()
#   <tag>_bench_{c4,c5,ref}.json          the bench lines (no profiler)
#   <tag>_launches.csv                    ncu launch list (gpu__time_duration) of a short bench run
#   <tag>_all_configs.txt                 all five configs end to end (tests/tools/config_times.py)
tag=${1:-r2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_pytest.txt
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1; tail -1 gpurun_out/${tag}_smoke.txt
python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err || { echo bench failed; tail -5 gpurun_out/${tag}_bench_c4.err; }
python bench.py --config c5 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err || tail -5 gpurun_out/${tag}_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
SHORT="python bench.py --steps 1 --warmup 3 --rays 2e8 --no-cpu-baseline"
$SHORT > gpurun_out/${tag}_plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches.csv $SHORT > gpurun_out/${tag}_ncu_launches.log 2>&1
python tests/tools/config_times.py 4 > gpurun_out/${tag}_all_configs.txt 2>&1; tail -8 gpurun_out/${tag}_all_configs.txt
for f in c4 c5 ref; do cut -c1-400 gpurun_out/${tag}_bench_$f.json; done
# steady-state ncu captures (16.7M-slot launches), one lane so that the launch index is deterministic
export VR_LANES=1
STEP="python scripts/profile_step.py 256e6 neutral"
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 12 -c 1 -f -o gpurun_out/prof_${tag}_trav $STEP > gpurun_out/${tag}_ncu_trav.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_${tag}_shade $STEP > gpurun_out/${tag}_ncu_shade.log 2>&1
STEP="python scripts/profile_step.py 256e6 ion"
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 12 -c 1 -f -o gpurun_out/prof_${tag}_shade_ion $STEP > gpurun_out/${tag}_ncu_shade_ion.log 2>&1
STEP="python scripts/profile_c5.py 2e8"
ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 8 -c 1 -f -o gpurun_out/prof_${tag}_trav_c5 $STEP > gpurun_out/${tag}_ncu_trav_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spreadKernel -s 8 -c 1 -f -o gpurun_out/prof_${tag}_spread_c5 $STEP > gpurun_out/${tag}_ncu_spread_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 8 -c 1 -f -o gpurun_out/prof_${tag}_shade_c5 $STEP > gpurun_out/${tag}_ncu_shade_c5.log 2>&1
unset VR_LANES
ls gpurun_out | grep ${tag} | wc -l
