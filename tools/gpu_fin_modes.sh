#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x 2>&1 | tail -2 > gpurun_out/l_fin.log
for m in 0 2; do echo "== RLG_FIN_DEBUG=$m"; RLG_FIN_DEBUG=$m timeout 120 python tools/step_breakdown.py 2>&1 | tail -1; done >> gpurun_out/l_fin.log 2>&1
timeout 120 python tools/step_breakdown.py 64 16384 16384 2>&1 | tail -1 >> gpurun_out/l_fin.log
