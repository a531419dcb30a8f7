#!/bin/bash
# ncu --set full of the training compositor (forward / backward) on the headline sample set; summaries made on the box
set -u
mkdir -p gpurun_out
python tools/run_kernel.py 18 > gpurun_out/run_kernel.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/run_kernel.log; exit 1; }
tail -1 gpurun_out/run_kernel.log
cap() {  # name kernel skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/run_kernel.py 18 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  python tools/summarize_profiles.py ncu gpurun_out/prof_$1.ncu-rep gpurun_out/r02h_ncu_$1.txt "ncu --set full --clock-control none, tools/run_kernel.py 18 (2^18 rays of the headline workload), final build of round 2" > /dev/null 2>&1
  python tools/ncu_lines.py gpurun_out/prof_$1.ncu-rep 45 > gpurun_out/r02h_ncu_$1_lines.txt 2>&1
  rm -f gpurun_out/prof_$1.ncu-rep
}
cap composite_train_fw composite_train_fw_kernel 1
cap composite_train_bw composite_train_bw_kernel 1
