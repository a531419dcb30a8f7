#!/bin/bash
# r01e evidence in one gpurun call: headline bench, playground bench, launch list of 2 timed steps, and ncu --set full
# captures of the street-shape scatter in both block orders (level chunk fastest / slowest).
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_e.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --workload playground --steps 10 --warmup 3 --pretrain 60 > gpurun_out/bench_playground_e3.log 2>&1; echo "playground rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --pretrain 0 --no-render"
timeout 300 $CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_e.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_e.csv
timeout 200 python tools/hash_order_probe.py street > gpurun_out/hash_order3.log 2>&1 || { echo "plain probe failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hashgrid_bw_params_kernel -s 1 -c 1 -f -o gpurun_out/prof_street_scatter_order0 python tools/hash_order_probe.py street > gpurun_out/ncu_o0.log 2>&1; echo "order0 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hashgrid_bw_params_kernel -s 8 -c 1 -f -o gpurun_out/prof_street_scatter_order1 python tools/hash_order_probe.py street > gpurun_out/ncu_o1.log 2>&1; echo "order1 rc=$?"
ls -la gpurun_out/*.ncu-rep
