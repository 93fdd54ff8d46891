#!/bin/bash
# usage (under gpurun): bash tools/sanitize.sh <tag>   -- compute-sanitizer over small shapes of every kernel family
T=$1
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for fam in chamfer encoder train; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py $fam > gpurun_out/${T}_san_${tool}_${fam}.log 2>&1
    echo "rc=$?" >> gpurun_out/${T}_san_${tool}_${fam}.log
  done
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|rc=" gpurun_out/${T}_san_*.log > gpurun_out/${T}_san_summary.log
echo done
