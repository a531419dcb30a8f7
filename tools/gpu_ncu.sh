#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/run_kernel.py 17 > gpurun_out/run_kernel.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/run_kernel.log; exit 1; }
tail -1 gpurun_out/run_kernel.log
for k in mlp_fw_kernel mlp_bw_kernel hashgrid_bw_params_kernel march_count_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/run_kernel.py 17 > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"; tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
