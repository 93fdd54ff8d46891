#!/bin/bash
# One GPU session: tests, breakdown, bench, launch list, ncu captures. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/a_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python tools/step_breakdown.py > gpurun_out/a_breakdown.log 2>&1
timeout 300 python tools/step_breakdown.py 64 16384 16384 >> gpurun_out/a_breakdown.log 2>&1
timeout 300 python tools/try_tc.py > gpurun_out/a_trytc.log 2>&1
timeout 900 python bench.py > gpurun_out/a_bench.log 2>&1
timeout 300 python tools/run_encoder.py 3 > gpurun_out/a_plain_enc.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_tc -c 2 -f -o gpurun_out/prof_enc_r1 python tools/run_encoder.py 3 > gpurun_out/a_ncu_enc.log 2>&1
timeout 300 python tools/run_chamfer.py 3 > gpurun_out/a_plain_ch.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"finalize|bwd" -s 6 -c 3 -f -o gpurun_out/prof_fin_r1 python tools/run_chamfer.py 3 > gpurun_out/a_ncu_fin.log 2>&1
nvidia-smi > gpurun_out/a_smi.log 2>&1
echo done
