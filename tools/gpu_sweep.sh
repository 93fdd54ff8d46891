#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep_tile.py > gpurun_out/p_sweep.log 2>&1
timeout 300 python tools/sweep_tile.py 16 16384 16384 >> gpurun_out/p_sweep.log 2>&1
