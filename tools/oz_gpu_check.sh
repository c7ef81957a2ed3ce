#!/bin/bash
# runs the INT8 substitution-update checks and timings of tools/oz_test on a GPU box; every case in its own process
# and under timeout, so that a trap or hang in one case cannot take the others (or the box) down
mkdir -p gpurun_out
LOG=gpurun_out/${1:-oz_test}.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 120 tools/oz_test "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run check 7 1024 256 7
run check 8 16384 64 127
run time 7 16384 9472 127 0 0 32
run time 6 16384 9472 127
run time 8 16384 9472 127
grep -E "RESULT|update_kernel|exit" $LOG
