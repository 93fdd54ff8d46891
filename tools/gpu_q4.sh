#!/bin/bash
T=${1:-w}
timeout 300 python tools/try_tcfilter.py > gpurun_out/${T}_try.log 2>&1; echo rc=$? >> gpurun_out/${T}_try.log
bash tools/gpu_tfdbg.sh > gpurun_out/${T}_dbg.log 2>&1
timeout 300 python tools/step_breakdown.py 2>&1 | grep tensor > gpurun_out/${T}_breakdown.log
timeout 300 python tools/step_breakdown.py 64 16384 16384 2>&1 | grep tensor >> gpurun_out/${T}_breakdown.log
timeout 300 python tools/step_breakdown.py 1024 2048 2048 2>&1 | grep tensor >> gpurun_out/${T}_breakdown.log
timeout 900 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
