#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vren_gpu.py -q -m gpu --timeout=300 > gpurun_out/test_vren.log 2>&1; echo "vren tests rc=$?"; tail -3 gpurun_out/test_vren.log
timeout 600 python bench.py --steps 10 --warmup 3 --pretrain 100 --no-render > gpurun_out/bench_quick.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.log'):
    if l.startswith('{"metric"'):
        d=json.loads(l)
        print('ms/step', round(d['ms_per_step'],3), 'Mrays/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'spr', round(d['config']['samples_per_ray'],1))
        for k,v in d['kernels'].items(): print(' ', k, round(v['ms'],3))
PY
timeout 300 python tools/step_profile.py 40 > gpurun_out/step_profile.txt 2>&1; grep -E "ngp::|Self CUDA time total" gpurun_out/step_profile.txt | cut -c1-75,150-230 | head -24
