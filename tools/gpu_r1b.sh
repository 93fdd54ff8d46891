#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/b_smoke.log 2>&1
timeout 1200 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x > gpurun_out/b_pytest_chamfer.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_chamfer.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_all.log
timeout 300 python tools/step_breakdown.py > gpurun_out/b_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 64 16384 16384 >> gpurun_out/b_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 128 2048 1400 >> gpurun_out/b_breakdown.log 2>&1
timeout 300 python tools/sweep_tile.py > gpurun_out/b_sweep.log 2>&1
timeout 300 python tools/sweep_tile.py 8 16384 16384 >> gpurun_out/b_sweep.log 2>&1
echo done
