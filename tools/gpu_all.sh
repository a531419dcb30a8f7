#!/bin/bash
# One gpurun call: the whole GPU suite, smoke, and a bench run.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/test_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:---steps 20 --warmup 5 --pretrain 400} > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log
