#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_density_net_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_dn.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/test_dn.log
echo "== probe"; timeout 900 python tools/occ_update_probe.py > gpurun_out/occ_update_probe.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/occ_update_probe.txt
echo "== scatter levels"; timeout 600 python tools/scatter_levels_probe.py > gpurun_out/scatter_levels_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/scatter_levels_probe.txt
echo "== scatter levels under ncu (request counts)"; NCU=1 timeout 900 ncu --metrics lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum,gpu__time_duration.sum --clock-control none -k regex:hashgrid_bw_params --csv --log-file gpurun_out/scatter_levels_ncu.csv python tools/scatter_levels_probe.py > gpurun_out/scatter_levels_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/scatter_levels_ncu.log
