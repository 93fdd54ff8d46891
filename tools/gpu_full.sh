#!/bin/bash
set -u
T=${1:-s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_all.log
timeout 300 python tools/step_breakdown.py > gpurun_out/${T}_breakdown.log 2>&1
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-encoder > gpurun_out/${T}_bench.log 2>&1
echo done
