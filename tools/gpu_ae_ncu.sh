#!/bin/bash
# usage (under gpurun): bash tools/gpu_ae_ncu.sh <tag> [B]   -- ncu launch list of eager AE training steps
T=$1; B=${2:-16}
mkdir -p gpurun_out
python tools/ae_eager.py $B 3 > gpurun_out/${T}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python tools/ae_eager.py $B 3 > gpurun_out/${T}_ncu.log 2>&1
echo done
