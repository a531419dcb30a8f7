#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --pretrain 0 --no-render"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python tools/run_kernel.py 18 > gpurun_out/run_kernel.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hashgrid_bw_params_kernel -s 1 -c 1 -f -o gpurun_out/prof_hashgrid_bw_v2 python tools/run_kernel.py 18 > gpurun_out/ncu_c.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hashgrid_fw_kernel -s 1 -c 1 -f -o gpurun_out/prof_hashgrid_fw_v2 python tools/run_kernel.py 18 > gpurun_out/ncu_d.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mlp_bw_kernel -s 2 -c 1 -f -o gpurun_out/prof_mlp_bw_rgb_v2 python tools/run_kernel.py 18 > gpurun_out/ncu_e.log 2>&1
echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
