#!/bin/bash
# quick Chamfer check: parity tests + timing breakdown + variant sweep
set -u
mkdir -p gpurun_out
T=${1:-c}
timeout 1200 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest_chamfer.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_chamfer.log
timeout 300 python tools/step_breakdown.py > gpurun_out/${T}_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 64 16384 16384 >> gpurun_out/${T}_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 128 2048 1400 >> gpurun_out/${T}_breakdown.log 2>&1
timeout 300 python tools/sweep_tile.py > gpurun_out/${T}_sweep.log 2>&1
echo done
