#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/n_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 200 -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/n_ncu.log 2>&1
timeout 300 python tools/run_chamfer.py 3 > gpurun_out/n_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chamfer_filter|finalize2|chamfer_bwd" -s 3 -c 3 -f -o gpurun_out/prof_chamfer_r1 python tools/run_chamfer.py 3 > gpurun_out/n_ncu2.log 2>&1
timeout 300 python tools/run_encoder.py 3 > gpurun_out/n_plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_tc -c 2 -f -o gpurun_out/prof_encoder_r1 python tools/run_encoder.py 3 > gpurun_out/n_ncu3.log 2>&1
echo done
