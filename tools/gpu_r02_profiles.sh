#!/bin/bash
# ncu --set full captures of the final round-2 kernels (one launch each, on the headline sample set / 2 M samples for the density net)
set -u
mkdir -p gpurun_out
python tools/run_kernel.py 18 > gpurun_out/run_kernel.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/run_kernel.log; exit 1; }
tail -1 gpurun_out/run_kernel.log
cap() {  # name kernel skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/run_kernel.py 18 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  # gpurun brings back at most 64 MiB: keep the text summaries (raw metrics + per-line profile), drop the report
  python tools/summarize_profiles.py ncu gpurun_out/prof_$1.ncu-rep gpurun_out/r02d_ncu_$1.txt "ncu --set full --clock-control none, tools/run_kernel.py 18 (2^18 rays of the headline workload), final build of round 2" > /dev/null 2>&1
  python tools/ncu_lines.py gpurun_out/prof_$1.ncu-rep 45 > gpurun_out/r02d_ncu_$1_lines.txt 2>&1
  rm -f gpurun_out/prof_$1.ncu-rep
}
cap hashgrid_bw_params_kernel hashgrid_bw_params_kernel 1
cap hashgrid_fw_kernel hashgrid_fw_kernel 1
cap mlp_sigma_fw mlp_fw_kernel 1
cap mlp_rgb_fw mlp_fw_kernel 3
cap mlp_sigma_bw mlp_bw_kernel 1
cap mlp_rgb_bw mlp_bw_kernel 3
cap march_count_kernel march_count_kernel 1
python tools/dn_ncu_driver.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:density_net -s 2 -c 2 -f -o gpurun_out/prof_density_net python tools/dn_ncu_driver.py > gpurun_out/ncu_density_net.log 2>&1; echo "density_net rc=$?"
ncu -i gpurun_out/prof_density_net.ncu-rep --page raw --csv > gpurun_out/r02d_ncu_density_net_raw.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/prof_density_net.ncu-rep 60 > gpurun_out/r02d_ncu_density_net_lines.txt 2>&1
rm -f gpurun_out/prof_density_net.ncu-rep
ls -la gpurun_out/ | tail -30
