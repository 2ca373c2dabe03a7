#!/bin/bash
# GPU session r03a: parity on the coprime split, bench A/B (gt_split 1/0), other shapes, tile-load ceiling, ncu.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r03a_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r03a_bench_gt.json 2> gpurun_out/r03a_bench_gt.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --opt gt_split=0 > gpurun_out/r03a_bench_ct.json 2> gpurun_out/r03a_bench_ct.err
timeout 300 python tools/bench_configs.py > gpurun_out/r03a_configs_gt.log 2>&1
timeout 300 python tools/bench_configs.py gt_split=0 61380 163680 > gpurun_out/r03a_configs_ct.log 2>&1
timeout 120 ./tools/microbench/tile_bw 32 > gpurun_out/r03a_tile_bw_32.log 2>&1
timeout 120 ./tools/microbench/tile_bw 512 > gpurun_out/r03a_tile_bw_512.log 2>&1
timeout 600 bash tools/gpu_profile.sh r03a > gpurun_out/r03a_profile.log 2>&1
ls -la gpurun_out
