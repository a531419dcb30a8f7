#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/run_kernel.py 17 > gpurun_out/run_kernel.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/run_kernel.log; exit 1; }
tail -1 gpurun_out/run_kernel.log
ncu --set full --clock-control none --import-source on -k regex:mlp_bw_kernel -s 2 -c 1 -f -o gpurun_out/prof_mlp_bw_rgb python tools/run_kernel.py 17 > gpurun_out/ncu_a.log 2>&1; echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mlp_bw_kernel -s 0 -c 1 -f -o gpurun_out/prof_mlp_bw_sigma python tools/run_kernel.py 17 > gpurun_out/ncu_b.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
