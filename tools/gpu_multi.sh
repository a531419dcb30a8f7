#!/bin/bash
# multi-GPU checks: NCCL equivalence test (2 ranks), then bench lines at N = $1 GPUs (weak lego, strong playground after / sharded, strong street)
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
echo "== dist test"; timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu --timeout=500 > gpurun_out/test_dist_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/test_dist_gpu.log
COMMON="--no-cpu --ref-steps 0 --no-other-configs"
echo "== lego weak N=$N"; timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 10 --pretrain 200 $COMMON > gpurun_out/bench_lego_weak_${N}gpu.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "rays/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_lego_weak_${N}gpu.log | head -3
for EX in after sharded; do
  echo "== playground strong N=$N exchange=$EX"; NGP_DP_EXCHANGE=$EX timeout 900 $TR bench.py --gpus $N --workload playground --scaling strong --steps 10 --warmup 3 --pretrain 30 $COMMON --no-render > gpurun_out/bench_playground_strong_${N}gpu_$EX.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "rays/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_playground_strong_${N}gpu_$EX.log | head -3
done
echo "== street strong N=$N"; timeout 900 $TR bench.py --gpus $N --workload street --scaling strong --steps 10 --warmup 3 --pretrain 30 $COMMON --no-render > gpurun_out/bench_street_strong_${N}gpu.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "rays/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_street_strong_${N}gpu.log | head -3
