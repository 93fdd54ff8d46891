#!/bin/bash
T=${1:-x}
timeout 300 python tools/step_breakdown.py 2>&1 | grep tensor > gpurun_out/${T}_breakdown.log
timeout 300 python tools/step_breakdown.py 64 16384 16384 2>&1 | grep tensor >> gpurun_out/${T}_breakdown.log
bash tools/gpu_tfdbg.sh 2>&1 | tail -2 >> gpurun_out/${T}_breakdown.log
timeout 300 python tools/try_tcfilter.py 2>&1 | grep -E "MATCH|MISMATCH" >> gpurun_out/${T}_breakdown.log
