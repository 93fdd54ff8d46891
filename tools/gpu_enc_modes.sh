#!/bin/bash
mkdir -p gpurun_out
for m in 16 23; do echo "== RLG_TC_DEBUG=$m"; RLG_TC_DEBUG=$m timeout 120 python tools/run_encoder.py 2 2>&1 | tail -30; done > gpurun_out/j_modes.log 2>&1
