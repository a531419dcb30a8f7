#!/bin/bash
# One gpurun call: mint goldens from the reference kernels, run the GPU parity suites, smoke, short bench.
# Every stage has its own timeout and log under gpurun_out/ so a hang in one does not lose the others.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== golden"; timeout 600 python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/golden.log
echo "== vren tests"; timeout 900 python -m pytest tests/test_vren_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_vren.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/test_vren.log
echo "== tcnn grid/sh tests"; timeout 600 python -m pytest tests/test_tcnn_gpu.py -q -m gpu -k "hashgrid or sh4" --timeout=300 > gpurun_out/test_grid.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/test_grid.log
echo "== tcnn mlp tests"; timeout 300 python -m pytest tests/test_tcnn_gpu.py -q -m gpu -k "mlp" --timeout=120 -x > gpurun_out/test_mlp.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/test_mlp.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --pretrain 40 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench.log
