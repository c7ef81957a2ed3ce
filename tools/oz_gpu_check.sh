#!/bin/bash
# runs the INT8 substitution-update checks of tools/oz_test on a GPU box; every case in its own process and under
# timeout, so that a trap or hang in one case cannot take the others (or the box) down
mkdir -p gpurun_out
LOG=gpurun_out/${1:-oz_test}.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 120 tools/oz_test "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG
run check 1 512 128 1 128 256
run check 1 512 128 1 256 128
run check 2 512 128 2 128 256
run check 7 1024 256 7 128 256
run check 7 1024 256 7 256 128
run check 6 1024 256 5 128 256
run check 8 1024 256 3 128 256
run check 7 16384 64 127 128 256
run time 7 16384 9472 127
run time 6 16384 9472 127
run time 8 16384 9472 127
tail -n 60 $LOG
