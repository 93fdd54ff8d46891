#!/bin/bash
# usage (under gpurun): bash tools/gpu_bwd_profile.sh <tag>  -- ncu evidence for the Chamfer backward kernels: plain runs
# first (must exit 0), then ncu --set full at cfg2 (inside a training step) and at the cfg5 shape (stand-alone calls,
# default and reproducible variants); summaries only come back (tools/export_profile.py)
T=$1
mkdir -p gpurun_out
python tools/run_chamfer.py 3 > gpurun_out/${T}_cfg2_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chamfer_bwd -s 1 -c 2 -f -o /tmp/${T}_cfg2 python tools/run_chamfer.py 3 > gpurun_out/${T}_cfg2_ncu.log 2>&1
python tools/export_profile.py /tmp/${T}_cfg2.ncu-rep gpurun_out/${T}_cfg2_ncu.txt > /dev/null 2>&1
python tools/bwd_probe.py 64 16384 16384 > gpurun_out/${T}_cfg5_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chamfer_bwd -s 2 -c 8 -f -o /tmp/${T}_cfg5 python tools/bwd_probe.py 64 16384 16384 > gpurun_out/${T}_cfg5_ncu.log 2>&1
python tools/export_profile.py /tmp/${T}_cfg5.ncu-rep gpurun_out/${T}_cfg5_ncu.txt > /dev/null 2>&1
ncu -i /tmp/${T}_cfg5.ncu-rep --page source --csv > /tmp/${T}_src.csv 2>/dev/null && python tools/ncu_stalls.py /tmp/${T}_src.csv 12 > gpurun_out/${T}_cfg5_stalls.txt 2>&1
echo done
