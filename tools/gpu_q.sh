#!/bin/bash
T=${1:-t}
timeout 300 python tools/step_breakdown.py > gpurun_out/${T}_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 64 16384 16384 >> gpurun_out/${T}_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 128 2048 1400 >> gpurun_out/${T}_breakdown.log 2>&1
bash tools/gpu_tfdbg.sh >> gpurun_out/${T}_breakdown.log 2>&1
