#!/bin/bash
for f in 1 0; do
RLG_CHAMFER_FUSE=$f timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-encoder > gpurun_out/ll_plain_$f.log 2>&1
RLG_CHAMFER_FUSE=$f timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 400 -c 120 --csv --log-file gpurun_out/ll_$f.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-encoder > gpurun_out/ll_ncu_$f.log 2>&1
done
