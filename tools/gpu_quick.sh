#!/bin/bash
# quick loop: selected GPU tests + kernel sweep + short bench
set -u
mkdir -p gpurun_out
echo "== tests ${TESTS:-tests/test_tcnn_gpu.py tests/test_model_gpu.py}"; timeout 900 python -m pytest ${TESTS:-tests/test_tcnn_gpu.py tests/test_model_gpu.py} -q -m gpu --timeout=300 -x > gpurun_out/test_quick.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/test_quick.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --pretrain 400 --no-render > gpurun_out/bench_quick.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.log'):
    if l.startswith('{"metric"'):
        d=json.loads(l)
        print('ms/step', round(d['ms_per_step'],3), 'Mrays/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'spr', round(d['config']['samples_per_ray'],2), 'psnr', round(d['config']['psnr_after_pretrain'],2))
        print({k: round(v['ms'],3) for k,v in d['kernels'].items()})
PY
tail -3 gpurun_out/bench_quick.log | cut -c1-300
