#!/bin/bash
set -u
T=${1:-h}
mkdir -p gpurun_out
timeout 120 python tools/try_tc.py > gpurun_out/${T}_trytc.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_trytc.log
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest_enc.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_enc.log
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/${T}_smi.log 2>&1
echo done
