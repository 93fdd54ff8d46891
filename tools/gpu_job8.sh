#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_job8.sh <tag> <N> <reserve> [<reserve> ...]  -- headline only, per reserved-SM setting
T=$1; N=$2; shift 2
mkdir -p gpurun_out
for R in "$@"; do
  RLG_RESERVED_SMS=$R timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 400 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/${T}_r${R}.log 2>&1
  echo "rc=$?" >> gpurun_out/${T}_r${R}.log
done
echo done
