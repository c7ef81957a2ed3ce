#!/bin/bash
# round 2, call 5 (1 GPU): dataflow TRSV (tests + timing), launch lists of C2 (GPC) and of a 1/8 shard of the C4 step
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call5
timeout 600 python -m pytest tests/test_gpu_trsv_flow.py -m gpu -x -q > $O.pytest_trsv.log 2>&1; echo "pytest trsv rc=$?"; tail -5 $O.pytest_trsv.log
timeout 300 python tools/trsv_probe.py 2048 5000 16384 50000 > $O.trsv_probe.log 2>&1; echo "probe rc=$?"; cat $O.trsv_probe.log
timeout 300 python tests/probes/bench_small.py C1 C2 C3 > $O.small.log 2>&1; echo "small rc=$?"; cat $O.small.log
python tests/probes/bench_small.py C2 > $O.c2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O.launches_c2.csv python tests/probes/bench_small.py C2 > $O.ncu_c2.log 2>&1
echo "c2 launch list rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-sample 0 --int8-peak-seconds 0.05 --test-size 125000"
$CMD > $O.plain_c4_shard.json 2> $O.plain_c4_shard.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O.launches_c4_shard.csv $CMD > $O.ncu_c4_shard.log 2>&1
echo "c4 shard launch list rc=$?"; wc -l $O.launches_c4_shard.csv $O.launches_c2.csv
