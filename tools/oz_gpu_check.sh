#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/${1:-oz_test}.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 120 tools/oz_test "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run check 7 1024 256 7 1 0 8
run time 7 16384 9472 127 1 0 8
run time 7 16384 9472 127 0 0 8
run time 7 16384 9472 127 1 0 10
grep -E "RESULT|update_kernel|exit|mismatch" $LOG | head -60
