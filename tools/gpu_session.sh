#!/bin/bash
# One GPU-box session (run under gpurun): parity suite, bench (all blocks), reference arm, other shapes,
# then the ncu captures of tools/gpu_profile.sh. Outputs land in gpurun_out/<tag>_*; the summaries that
# are meant to be kept are copied to profiles/ by hand. Multi-GPU: `gpurun --gpus N -- bash tools/gpu_session.sh TAG N`.
TAG=${1:-session}
N=${2:-1}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
if [ "$N" -gt 1 ]; then
  timeout 300 python tests/test_nccl_abi.py $N > gpurun_out/${TAG}_nccl_abi_${N}.log 2>&1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  exit 0
fi
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 300 python tools/bench_configs.py > gpurun_out/${TAG}_configs.log 2>&1
timeout 600 bash tools/gpu_profile.sh ${TAG} > gpurun_out/${TAG}_profile.log 2>&1
