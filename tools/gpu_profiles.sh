#!/bin/bash
# usage (under gpurun): bash tools/gpu_profiles.sh <tag>   -- the round's ncu evidence: every command is first run plain
# (must exit 0), then under ncu; the reports are summarised on the box (tools/export_profile.py) and only the summaries
# (and the Chamfer report, for source-level stall samples) come back: gpurun_out/ is capped at 64 MiB
T=$1
mkdir -p gpurun_out
run_full() {  # name, keep-report(0/1), kernel regex, skip, count, command...
  local name=$1 keep=$2 k=$3 s=$4 c=$5; shift 5
  "$@" > gpurun_out/${T}_${name}_plain.log 2>&1 || { echo "plain run failed: $name"; return; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/${T}_${name} "$@" > gpurun_out/${T}_${name}_ncu.log 2>&1
  echo "rc=$?" >> gpurun_out/${T}_${name}_ncu.log
  python tools/export_profile.py /tmp/${T}_${name}.ncu-rep gpurun_out/${T}_${name}_ncu.txt > /dev/null 2>&1
  if [ "$keep" = "1" ]; then cp /tmp/${T}_${name}.ncu-rep gpurun_out/; fi
}
run_full chamfer 1 "chamfer_tcsweep|chamfer_bwd" 4 2 python tools/run_chamfer.py 3
run_full enc_bf16 0 "encoder_tc_kernel" 1 1 python tools/run_encoder.py 3 256 2048 bf16
run_full enc_cfg 0 "encoder_layer" 4 4 python tools/run_encoder_cfg.py 2 256 2048 fp32x
run_full train 0 "encoder_wgrad|encoder_layer_kernel|bn_|layer0_|wgrad_reduce|pool_|train_" 45 45 python tools/ae_eager.py 16 3
# launch lists (device time per launch): the headline bench with graph nodes, eager AE steps, an eager environment step
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 200 -c 300 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_ae_launches.csv python tools/ae_eager.py 16 3 > gpurun_out/${T}_ae_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_env_launches.csv python tools/env_eager.py 1024 > gpurun_out/${T}_env_ncu.log 2>&1
du -sh gpurun_out
echo done
