#!/bin/bash
# GPU session r03j: full parity suite, bench (configs + sweep blocks), other shapes
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r03j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r03j_bench.json 2> gpurun_out/r03j_bench.err
timeout 300 python tools/bench_configs.py > gpurun_out/r03j_configs.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03j_bench_ref.json 2> gpurun_out/r03j_bench_ref.err
