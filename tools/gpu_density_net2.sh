#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_density_net_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_dn.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/test_dn.log
echo "== probe"; timeout 900 python tools/occ_update_probe.py > gpurun_out/occ_update_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/occ_update_probe.txt | tail -16
echo "== ncu"; timeout 900 ncu --set full --import-source on --clock-control none -k regex:density_net -c 4 -o gpurun_out/dn_probe -f python tools/dn_ncu_driver.py > gpurun_out/ncu_dn.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_dn.log
