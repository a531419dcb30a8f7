#!/bin/bash
# One gpurun call that regenerates the evidence under profiles/: launch list of 2 timed steps (ncu
# gpu__time_duration) + ncu --set full captures of the hash-grid kernels on the step's real sample set.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --pretrain 0 --no-render"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python tools/run_kernel.py 18 > gpurun_out/run_kernel.log 2>&1 || { echo "plain run failed"; exit 1; }
for k in hashgrid_bw_params_kernel hashgrid_fw_kernel march_count_kernel composite_train_bw_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/run_kernel.py 18 > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
