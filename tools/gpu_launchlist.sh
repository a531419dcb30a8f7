#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== scatter levels"; timeout 600 python tools/scatter_levels_probe.py > gpurun_out/scatter_levels_probe.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/scatter_levels_probe.txt
CMD="python bench.py --steps 2 --warmup 3 --pretrain 100 --no-render --no-cpu --ref-steps 0 --no-other-configs"
$CMD > gpurun_out/plain_bench.log 2>&1; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
