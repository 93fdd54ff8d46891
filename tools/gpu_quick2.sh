#!/bin/bash
set -u
T=${1:-o}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_chamfer_gpu.py tests/test_dropin_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python tools/step_breakdown.py > gpurun_out/${T}_breakdown.log 2>&1
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-encoder > gpurun_out/${T}_bench.log 2>&1
echo done
