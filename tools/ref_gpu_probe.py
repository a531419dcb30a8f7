"""Probe: cost / memory / PSNR of the reference-GPU arm (baseline/ref_train.py) at a given ray batch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from baseline import ref_train
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
field = sys.argv[3] if len(sys.argv) > 3 else "ngp_pl"
vren = sys.argv[4] if len(sys.argv) > 4 else "ref"
tcnn = sys.argv[5] if len(sys.argv) > 5 else "standin"
out = ref_train.run_reference_gpu(rays=rays, steps_total=steps, timed_last=min(5, steps // 2), field=field, vren=vren, tcnn=tcnn)
print(json.dumps(out))
