#!/bin/bash
# runs INT8 substitution-update checks / timings of tools/oz_test on a GPU box; every case in its own process and
# under timeout, so that a trap or hang in one case cannot take the others (or the box) down
mkdir -p gpurun_out
LOG=gpurun_out/${1:-oz_test}.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 100 tools/oz_test "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run check 7 1024 256 7 2 0
run check 6 1024 256 5 2 0
run check 8 16384 128 127 2 0
run time 7 16384 18944 127 2 0
run time 7 16384 18944 127 0 0
grep -E "RESULT|update_kernel|exit|mismatch" $LOG | head -40
