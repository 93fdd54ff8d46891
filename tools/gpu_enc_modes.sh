#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 3 4 7; do echo "== RLG_TC_DEBUG=$m"; RLG_TC_DEBUG=$m timeout 120 python tools/try_tc.py 2>&1 | tail -1; done > gpurun_out/j_modes.log 2>&1
