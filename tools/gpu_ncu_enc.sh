#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_encoder.py 3 > gpurun_out/i_plain_enc.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_tc -c 2 -f -o gpurun_out/prof_enc2_r1 python tools/run_encoder.py 3 > gpurun_out/i_ncu_enc.log 2>&1
echo done
