#!/bin/bash
timeout 300 python tools/try_tcfilter.py > gpurun_out/tc5.log 2>&1; echo rc=$? >> gpurun_out/tc5.log
bash tools/gpu_tfdbg.sh >> gpurun_out/tc5.log 2>&1
