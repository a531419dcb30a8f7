#!/bin/bash
# round 2, late: x-major bitfield in the training marcher, group kernels (compositors / distortion / Ref-NeRF losses) walking 8 rays per warp
set -u
mkdir -p gpurun_out
echo "== parity tests (marcher / compositor / losses, incl. the 2^18-ray cases vs the live reference kernels)"
timeout 900 python -m pytest tests/test_vren_gpu.py tests/test_full_size_gpu.py -q -m gpu --timeout=300 -x > gpurun_out/h_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/h_tests.log
echo "== compositor / distortion sweep"
timeout 600 python tools/composite_sweep.py > gpurun_out/h_composite_sweep.txt 2>&1; echo "rc=$?"; cat gpurun_out/h_composite_sweep.txt | tail -8
echo "== marcher count pass"
for cfg in "NGP_MARCH_LINEAR=0" "NGP_MARCH_LINEAR=1" "NGP_MARCH_LINEAR=1 NGP_MARCH_CTAS_PER_SM=2" "NGP_MARCH_LINEAR=1 NGP_MARCH_CTAS_PER_SM=3" "NGP_MARCH_LINEAR=1 NGP_MARCH_CTAS_PER_SM=6"; do
  echo -n "$cfg  "; env $cfg timeout 300 python tools/march_count_probe.py 2>&1 | tail -1
done | tee gpurun_out/h_march_probe.txt
B="python bench.py --steps 30 --warmup 10 --pretrain 400 --no-render --no-cpu --ref-steps 0 --no-other-configs"
echo "== quick bench, new defaults"; timeout 600 $B > gpurun_out/h_bench_new.log 2>&1; echo "rc=$?"
echo "== quick bench, previous behaviour"; NGP_MARCH_LINEAR=0 NGP_COMPOSITE_TILED=0 NGP_COMPOSITE_BLOCK=256 timeout 600 $B > gpurun_out/h_bench_old.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ('h_bench_new', 'h_bench_old'):
    for l in open(f'gpurun_out/{f}.log'):
        if l.startswith('{"metric"'):
            d = json.loads(l)
            print(f, 'ms/step', round(d['ms_per_step'], 3), 'Mrays/s', round(d['value'] / 1e6, 2), 'e2e', round(d['e2e']['value'] / 1e6, 2), 'spr', round(d['config']['samples_per_ray'], 2),
                  'psnr', round(d['config']['psnr_after_pretrain'], 2))
PY
