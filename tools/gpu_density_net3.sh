#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_density_net_gpu.py tests/test_model_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_dn.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_dn.log
echo "== probe"; timeout 900 python tools/occ_update_probe.py > gpurun_out/occ_update_probe.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/occ_update_probe.txt
echo "== playground bench"; timeout 900 python bench.py --workload playground --steps 20 --warmup 5 --pretrain 60 --no-render --no-cpu > gpurun_out/bench_playground.log 2>&1; echo "rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_playground.log
echo "== playground bench, TF32 torch-GEMM density net"; NGP_DENSITY_NET_TC=0 timeout 900 python bench.py --workload playground --steps 20 --warmup 5 --pretrain 60 --no-render --no-cpu > gpurun_out/bench_playground_tf32.log 2>&1; echo "rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"psnr_after_pretrain": [0-9.]*' gpurun_out/bench_playground_tf32.log gpurun_out/bench_playground.log
