#!/bin/bash
mkdir -p gpurun_out
SWEEP_VARIANTS=${1:-2,10,11,12,13,14,15} timeout 300 python tools/sweep_tile.py > gpurun_out/ko_sweep.log 2>&1
SWEEP_VARIANTS=${1:-2,10,11,12,13,14,15} timeout 300 python tools/sweep_tile.py 16 16384 16384 >> gpurun_out/ko_sweep.log 2>&1
