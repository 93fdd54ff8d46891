#!/bin/bash
T=${1:-u}
timeout 300 python tools/try_tcfilter.py > gpurun_out/${T}_try.log 2>&1; echo rc=$? >> gpurun_out/${T}_try.log
timeout 600 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python tools/step_breakdown.py > gpurun_out/${T}_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 64 16384 16384 >> gpurun_out/${T}_breakdown.log 2>&1
