#!/bin/bash
# ncu --set full captures of MLP launches of tools/run_kernel.py: $@ = subset of {sigma_fw rgb_fw sigma_bw rgb_bw}
set -u
mkdir -p gpurun_out
python tools/run_kernel.py 18 > gpurun_out/run_kernel.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/run_kernel.log; exit 1; }
tail -1 gpurun_out/run_kernel.log
cap() {  # name kernel skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/run_kernel.py 18 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
for w in "$@"; do
  case $w in
    sigma_fw) cap mlp_sigma_fw mlp_fw_kernel 1;;
    rgb_fw) cap mlp_rgb_fw mlp_fw_kernel 3;;
    sigma_bw) cap mlp_sigma_bw mlp_bw_kernel 1;;
    rgb_bw) cap mlp_rgb_bw mlp_bw_kernel 3;;
  esac
done
