#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the correlate kernels.
# Outputs land in gpurun_out/ (merged back by gpurun); summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_corr -s 40 -c 2 -f \
    -o gpurun_out/prof_corr_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
