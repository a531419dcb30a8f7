#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4 6; do echo "ctas/sm=$c"; NGP_MARCH_CTAS_PER_SM=$c python tools/compare_ref.py 18 2>&1 | grep -E "raymarching_train"; done
timeout 600 python -m pytest tests/test_vren_gpu.py -q -m gpu -k "march" --timeout=300 2>&1 | tail -2
