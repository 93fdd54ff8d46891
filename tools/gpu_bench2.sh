#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 600 --warmup 20 > gpurun_out/m_bench1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 600 --warmup 20 > gpurun_out/m_bench2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/m_ref2.log 2>&1
echo done
