#!/bin/bash
# density net on tcgen05: parity tests, the model tests around it, the playground-shaped bench and its step profile
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_density_net_gpu.py tests/test_model_gpu.py -q -m gpu --timeout=300 -x > gpurun_out/test_dn.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/test_dn.log
echo "== playground bench"; timeout 900 python bench.py --workload playground --steps 10 --warmup 3 --pretrain 100 --no-render --no-cpu > gpurun_out/bench_playground.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_playground.log | cut -c1-600
echo "== step profile"; timeout 600 python tools/step_profile_ngp.py > gpurun_out/step_profile_playground.txt 2>&1; echo "rc=$?"; head -45 gpurun_out/step_profile_playground.txt | cut -c1-120,180-260
