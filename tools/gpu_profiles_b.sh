#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-encoder > gpurun_out/pb_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 200 -c 300 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-encoder > gpurun_out/pb_ncu.log 2>&1
timeout 300 python tools/run_chamfer.py 3 > gpurun_out/pb_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chamfer_tcfilter|finalize2|chamfer_bwd" -s 3 -c 3 -f -o gpurun_out/prof_chamfer_r1b python tools/run_chamfer.py 3 > gpurun_out/pb_ncu2.log 2>&1
build/ubench2 > gpurun_out/ubench2.log 2>&1
echo done
