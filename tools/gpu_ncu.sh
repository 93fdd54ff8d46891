#!/bin/bash
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> <kernel-regex> <skip> <count> <command...>
# plain run first (must exit 0), then one ncu --set full capture of the matching kernels
set -u
T=$1; K=$2; S=$3; C=$4; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/${T}_prof "$@" > gpurun_out/${T}_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_ncu.log
