#!/bin/bash
# GPU session r03e: copy-engine-fed pair: parity, A/B sweep, bench, ncu.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r03e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03e_pytest.log
timeout 600 python tools/ab_v3.py quick > gpurun_out/r03e_ab_v3.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r03e_bench.json 2> gpurun_out/r03e_bench.err
timeout 600 bash tools/gpu_profile.sh r03e > gpurun_out/r03e_profile.log 2>&1
ls -la gpurun_out | tail -12
