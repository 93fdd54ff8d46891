timeout 300 python tools/step_breakdown.py 64 16384 16384 2>&1 | grep tensor > gpurun_out/y_breakdown.log
RLG_TF_TOP3=1 timeout 300 python tools/step_breakdown.py 2>&1 | grep tensor >> gpurun_out/y_breakdown.log
RLG_TF_TOP3=1 timeout 300 python tools/step_breakdown.py 1024 2048 2048 2>&1 | grep tensor >> gpurun_out/y_breakdown.log
timeout 300 python tools/step_breakdown.py 1024 2048 2048 2>&1 | grep tensor >> gpurun_out/y_breakdown.log
timeout 900 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x 2>&1 | tail -2 >> gpurun_out/y_breakdown.log
