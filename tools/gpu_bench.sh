#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --pretrain 400 --no-render > gpurun_out/bench_run$i.log 2>&1; echo "rc=$?"
python - $i <<'PY'
import json,sys
for l in open(f'gpurun_out/bench_run{sys.argv[1]}.log'):
    if l.startswith('{"metric"'):
        d=json.loads(l)
        print('ms/step', round(d['ms_per_step'],3), 'Mrays/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['config']['samples_per_ray_timed_steps'], 'psnr', round(d['config']['psnr_after_pretrain'],2))
PY
done
