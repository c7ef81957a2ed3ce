#!/bin/bash
# timing diagnostics of the INT8 update kernel (dbg bits: see UpdateParams)
mkdir -p gpurun_out
LOG=gpurun_out/${1:-oz_time}.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 120 tools/oz_test "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
for dbg in 32 33 34; do run time 7 16384 9472 127 128 256 $dbg; done
run time 7 16384 64 127 128 256 32
cat $LOG
