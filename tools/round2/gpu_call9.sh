#!/bin/bash
# round 2, call 9 (2 GPUs): dist tests with the new defaults, C4 bench on 2 GPUs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call9
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > $O.pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -5 $O.pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > $O.bench_2gpu.json 2> $O.bench_2gpu.err; echo "bench2 rc=$?"; tail -c 300 $O.bench_2gpu.err; grep "^{" $O.bench_2gpu.json | head -c 600
