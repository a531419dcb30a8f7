#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_tcnn_gpu.py tests/test_model_gpu.py tests/test_density_net_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_mlp.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/test_mlp.log
echo "== playground bench"; timeout 900 python bench.py --workload playground --steps 20 --warmup 5 --pretrain 60 --no-render --no-cpu > gpurun_out/bench_playground.log 2>&1; echo "rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"psnr_after_pretrain": [0-9.]*' gpurun_out/bench_playground.log
echo "== step profile"; timeout 600 python tools/step_profile_ngp.py > gpurun_out/step_profile_playground.txt 2>&1; echo "rc=$?"; grep -E 'mlp_|density_net|Self CUDA time total' gpurun_out/step_profile_playground.txt | cut -c1-100,170-260
echo "== lego bench"; timeout 600 python bench.py --steps 20 --warmup 5 --pretrain 400 --no-render --no-cpu --ref-steps 0 --no-other-configs > gpurun_out/bench_quick.log 2>&1; echo "rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_quick.log | head -2
