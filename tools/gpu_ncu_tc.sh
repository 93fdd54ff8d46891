#!/bin/bash
set -u
mkdir -p gpurun_out
export RLG_CHAMFER_SWEEP=tensor
timeout 300 python tools/run_chamfer.py 3 > gpurun_out/tcn_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chamfer_tcfilter" -s 1 -c 1 -f -o gpurun_out/prof_tcfilter python tools/run_chamfer.py 3 > gpurun_out/tcn_ncu.log 2>&1
echo done
