#!/bin/bash
cd "$(dirname "$0")/.."
python bench.py --cpu-seconds 8 > gpurun_out/r2s_bench_c4.json 2> gpurun_out/r2s_bench_c4.err || tail -20 gpurun_out/r2s_bench_c4.err
cut -c1-1200 gpurun_out/r2s_bench_c4.json
python bench.py --config c5 --cpu-seconds 8 > gpurun_out/r2s_bench_c5.json 2> gpurun_out/r2s_bench_c5.err || tail -20 gpurun_out/r2s_bench_c5.err
cut -c1-1200 gpurun_out/r2s_bench_c5.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err || tail -20 gpurun_out/r2s_bench_ref.err
cut -c1-600 gpurun_out/r2s_bench_ref.json
