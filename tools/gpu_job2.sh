#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_job2.sh <tag> <N> [steps]   -- the driver's multi-GPU bench launch
T=$1; N=$2; STEPS=${3:-400}
mkdir -p gpurun_out
timeout ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 20 --no-cpu-baseline > gpurun_out/${T}_bench${N}.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_bench${N}.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.log 2>&1
echo done
