#!/bin/bash
# GPU session r03g: copy-engine-fed pair: parity, A/B sweep, bench, ncu.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r03g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03g_pytest.log
timeout 300 python tools/bench_configs.py > gpurun_out/r03g_configs.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err
timeout 600 bash tools/gpu_profile.sh r03g > gpurun_out/r03g_profile.log 2>&1
ls -la gpurun_out | tail -12
