#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu captures of the correlate kernels.
# Outputs land in gpurun_out/ (merged back by gpurun); summaries are copied to profiles/ by hand.
# Kernels are profiled serialised (overlap_chunks=0) so each launch is one kernel on an idle GPU.
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_corr -s 80 -c 2 -f \
    -o gpurun_out/prof_corr_${TAG} $B --opt overlap_chunks=0 > gpurun_out/ncu_full_${TAG}.log 2>&1
# DRAM traffic with warm caches (no flush between replays): what the kernels really pull from HBM
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --cache-control none \
    --clock-control none -k regex:k_corr -s 80 -c 8 --csv --log-file gpurun_out/traffic_${TAG}.csv \
    $B --opt overlap_chunks=0 > /dev/null 2>&1
ls -la gpurun_out
