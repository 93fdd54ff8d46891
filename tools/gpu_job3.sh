#!/bin/bash
# usage (under gpurun): bash tools/gpu_job3.sh <tag>  -- step launch list of the headline bench (first 80 launches: eager warm-up steps + captured steps)
T=$1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 0 -c 80 --csv --log-file gpurun_out/${T}_step_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_step_ncu.log 2>&1
echo done
