timeout 300 python tools/dbg_hostgraph.py > gpurun_out/dbg.log 2>&1
for i in 1 2 3; do timeout 300 python -m pytest tests/test_pipeline_gpu.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -3 >> gpurun_out/dbg.log; done
