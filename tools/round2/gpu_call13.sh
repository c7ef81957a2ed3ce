#!/bin/bash
# round 2, call 13 (2 GPUs): BASELINE config 3 (fit() multi-start, rational quadratic, n = 5000, 16 starts) on 1 and 2 GPUs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call13
timeout 400 python tests/probes/bench_c3.py > $O.c3_1gpu.json 2> $O.c3_1gpu.err; echo "c3 1 GPU rc=$?"; cat $O.c3_1gpu.json | tail -1 | head -c 700; echo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tests/probes/bench_c3.py > $O.c3_2gpu.json 2> $O.c3_2gpu.err; echo "c3 2 GPUs rc=$?"; grep "^{" $O.c3_2gpu.json | head -c 700; echo
