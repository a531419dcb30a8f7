#!/bin/bash
# launch list of the headline step: every kernel's gpu__time_duration over the first 6000 launches of a short bench run
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --pretrain 100 --no-render --no-cpu --ref-steps 0 --no-other-configs"
$CMD > gpurun_out/plain_bench.log 2>&1; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
