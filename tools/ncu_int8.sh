#!/bin/bash
# ncu evidence for the INT8 variance pass (run under gpurun, one GPU):
#   launch list (device time of every kernel) of a reduced bench command: n = 16 384, m = 151 552 (2 chunks),
#   one warm-up step + one timed step (identical), all launches captured
#   (the --set full capture of oz::update_kernel<7> is taken from `tools/oz_test time 7 16384 9472 127`)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --parity-sample 0 --train-size 16384 --test-size 151552"
$B > gpurun_out/ncu_int8_plain.json 2> gpurun_out/ncu_int8_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_int8.csv $B > gpurun_out/ncu_int8_launches.log 2>&1
echo "launch list rc=$?"
