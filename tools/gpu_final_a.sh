#!/bin/bash
set -u
T=${1:-fa}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_all.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ref.log
echo done
