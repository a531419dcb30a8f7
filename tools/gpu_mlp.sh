#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== mlp tests"; timeout 300 python -m pytest tests/test_tcnn_gpu.py -q -m gpu -k "mlp" --timeout=120 -x > gpurun_out/test_mlp.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/test_mlp.log
echo "== sweep"; timeout 600 python tools/mlp_sweep.py > gpurun_out/mlp_sweep.log 2>&1; echo "rc=$?"; cat gpurun_out/mlp_sweep.log | tail -20
