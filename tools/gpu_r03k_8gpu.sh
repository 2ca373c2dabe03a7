#!/bin/bash
# 8-GPU session: NCCL inside the C ABI at 4 and 8 ranks, bench weak + strong + sweep at 8 and 4 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r03k_gpus.txt
timeout 300 python tests/test_nccl_abi.py 8 > gpurun_out/r03k_nccl_abi_8.log 2>&1; echo "rc=$?" >> gpurun_out/r03k_nccl_abi_8.log
timeout 300 python tests/test_nccl_abi.py 4 > gpurun_out/r03k_nccl_abi_4.log 2>&1; echo "rc=$?" >> gpurun_out/r03k_nccl_abi_4.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r03k_bench_${n}gpu.json 2> gpurun_out/r03k_bench_${n}gpu.err
done
