#!/bin/bash
# 2-GPU session: NCCL inside the C ABI, bench weak + strong scaling under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r03h_gpus.txt
timeout 300 python tests/test_nccl_abi.py 2 > gpurun_out/r03h_nccl_abi.log 2>&1; echo "rc=$?" >> gpurun_out/r03h_nccl_abi.log
timeout 300 python -m pytest tests/test_nccl_abi.py -m gpu -q > gpurun_out/r03h_nccl_pytest.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03h_bench_2gpu.json 2> gpurun_out/r03h_bench_2gpu.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r03h_bench_1gpu.json 2> gpurun_out/r03h_bench_1gpu.err
