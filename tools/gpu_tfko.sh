#!/bin/bash
for ko in 0 1 2 3; do
echo "== RLG_TF_KO=$ko"
RLG_TF_KO=$ko timeout 120 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.argv=['x']
import runpy
m = runpy.run_path('tools/try_tcfilter.py', run_name='notmain')
m['time_sweep'](32, 2048, 2048, True)
PY
done
