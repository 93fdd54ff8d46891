#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_chamfer.py 3 > gpurun_out/d_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chamfer_filter|finalize2|chamfer_bwd" -s 3 -c 3 -f -o gpurun_out/prof_filter_r1 python tools/run_chamfer.py 3 > gpurun_out/d_ncu.log 2>&1
echo done
