#!/bin/bash
# usage (under gpurun): bash tools/gpu_job.sh <tag> <step> [<step> ...]   -- each step writes gpurun_out/<tag>_<step>.log
set -u
T=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    chamfer)   timeout 1500 python -m pytest tests/test_chamfer_gpu.py -m gpu -q --maxfail=25 > gpurun_out/${T}_chamfer.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_chamfer.log ;;
    pytest)    timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log ;;
    breakdown) RLG_EXPERIMENTS_LIB=1 timeout 600 python tools/step_breakdown.py > gpurun_out/${T}_breakdown.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_breakdown.log ;;
    breakdown5) RLG_EXPERIMENTS_LIB=1 timeout 600 python tools/step_breakdown.py 8 16384 16384 > gpurun_out/${T}_breakdown5.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_breakdown5.log ;;
    smoke)     timeout 600 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log ;;
    bench)     timeout 1200 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log ;;
    benchq)    timeout 900 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/${T}_benchq.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_benchq.log ;;
    benchref)  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_benchref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_benchref.log ;;
    variants)  for so in gan-rl_3d_b200/lib/librlg_b200_exp_*.so; do echo "== $so" >> gpurun_out/${T}_variants.log; RLG_EXPERIMENTS_LIB=$PWD/$so timeout 300 python tools/step_breakdown.py 32 2048 2048 sphere tensor >> gpurun_out/${T}_variants.log 2>&1; done ;;
    encoder)   timeout 1500 python -m pytest tests/test_encoder_gpu.py -m gpu -q --maxfail=60 > gpurun_out/${T}_encoder.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_encoder.log ;;
    encprobe)  timeout 900 python tools/encoder_probe.py > gpurun_out/${T}_encprobe.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_encprobe.log ;;
    train)     timeout 900 python -m pytest tests/test_encoder_train_gpu.py -m gpu -q -s --maxfail=60 > gpurun_out/${T}_train.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_train.log ;;
    trainprobe) timeout 600 python tools/train_probe.py > gpurun_out/${T}_trainprobe.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_trainprobe.log ;;
    aestep)    timeout 900 python -m pytest tests/test_ae_step_gpu.py -m gpu -q -s > gpurun_out/${T}_aestep.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_aestep.log ;;
    envtest)   timeout 900 python -m pytest tests/test_environment_gpu.py -m gpu -q -s > gpurun_out/${T}_envtest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_envtest.log ;;
    property)  timeout 900 python -m pytest tests/test_chamfer_property_gpu.py -m gpu -q > gpurun_out/${T}_property.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_property.log ;;
    data)      timeout 900 python -m pytest tests/test_data_gpu.py tests/test_dropin_gpu.py -m gpu -q > gpurun_out/${T}_data.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_data.log ;;
    bwddet)    timeout 600 python -m pytest tests/test_chamfer_bwd_gpu.py -m gpu -q > gpurun_out/${T}_bwddet.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bwddet.log ;;
    bwdprobe)  timeout 600 python tools/bwd_probe.py > gpurun_out/${T}_bwdprobe.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bwdprobe.log
               for so in gan-rl_3d_b200/lib/librlg_b200_exp_bwd*.so; do RLG_EXPERIMENTS_LIB=$PWD/$so timeout 600 python tools/bwd_probe.py >> gpurun_out/${T}_bwdprobe.log 2>&1; done ;;
    steplist)  bash tools/gpu_job3.sh ${T} ;;
    *) echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.log 2>&1
echo done
