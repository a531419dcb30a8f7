#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== vren tests"; timeout 900 python -m pytest tests/test_vren_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_vren.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_vren.log
echo "== tcnn tests"; timeout 900 python -m pytest tests/test_tcnn_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_tcnn.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/test_tcnn.log
echo "== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py tests/test_optim_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_model.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/test_model.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --pretrain 300 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench.log | cut -c1-3000
echo "== step profile"; timeout 300 python tools/step_profile.py 40 > gpurun_out/step_profile.txt 2>&1; echo "rc=$?"; head -60 gpurun_out/step_profile.txt | cut -c1-200
