#!/bin/bash
# The first GPU calls of the next round, in order.  Each line is one gpurun call.
#
# 0. the non-gating tier (tests written without GPU time): python -m pytest tests -m gpu_next; promote green tests
/usr/local/graft/bin/gpurun --timeout 200 -- \
  'timeout 180 python -m pytest tests -m gpu_next -v > gpurun_out/pytest_gpu_next.log 2>&1; tail -8 gpurun_out/pytest_gpu_next.log'
# 1. the wide INT8 kernel at full C4 size (default run of the bench with --int8-tile 128); promote it to the default
#    (opt_int8_tile = 128 in csrc/common.cuh, --int8-tile default in bench.py) if `parity` is green and `value` beats
#    profiles/r1_bench_c4_int8_default.json (34.31 s):
/usr/local/graft/bin/gpurun --timeout 420 -- \
  'timeout 400 python bench.py --int8-tile 128 > gpurun_out/bench_c4_tile128.json 2> gpurun_out/bench_c4_tile128.err; tail -c 400 gpurun_out/bench_c4_tile128.err; head -c 900 gpurun_out/bench_c4_tile128.json'
# 2. the full GPU tier on the then-current default:
/usr/local/graft/bin/gpurun --timeout 300 -- \
  'timeout 280 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log'
# 3. 8 GPUs, C4 (charged 8x): scaling of train (distributed Cholesky) + INT8 predict per shard
/usr/local/graft/bin/gpurun --gpus 8 --timeout 300 -- \
  'timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 1 --warmup 3 > gpurun_out/bench_c4_8gpu_int8.json 2> gpurun_out/bench_c4_8gpu_int8.err; head -c 900 gpurun_out/bench_c4_8gpu_int8.json'
# 4. the CTA-pair kernel draft (csrc/ozaki_pair.cuh, not yet run on hardware): correctness against the digit emulation,
#    then its rate next to the wide and the default kernel (j = pair index; block rows 2j, 2j+1 against k < 256 j)
/usr/local/graft/bin/gpurun --timeout 300 -- \
  'for a in "check 7 1024 256 1 3 0" "check 7 1024 256 3 3 0" "check 8 16384 128 63 3 0" "time 7 16384 18944 63 3 0" "time 7 16384 18944 127 2 0" "time 7 16384 18944 127 0 0"; do timeout 100 tools/oz_test $a 2>&1 | grep -E "RESULT|update_kernel|mismatch|error"; done'
