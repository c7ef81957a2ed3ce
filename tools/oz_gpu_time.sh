#!/bin/bash
# timing diagnostics of the INT8 update kernel: full / feed only / MMA only / per-digit copies
mkdir -p gpurun_out
LOG=gpurun_out/${1:-oz_time}.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 120 tools/oz_test "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
for dbg in 0 1 2 4; do run time 7 16384 9472 127 128 256 $dbg; done
run time 7 16384 18944 127 128 256 0
run time 8 16384 9472 127 128 256 2
run time 6 16384 9472 127 128 256 2
grep -E "update_kernel|exit" $LOG
