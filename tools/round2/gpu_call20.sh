#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P=gaussian-process-regression_b200
gcc -std=c11 -Wall -D_GNU_SOURCE -Itests/stubs -Iinclude $P/src/gprc_shim.c tests/stubs/mini_r.c tests/shim_exec.c -L$P -lgprc -lm -o /tmp/shim_exec && LD_LIBRARY_PATH=$P /tmp/shim_exec > gpurun_out/r2_call20.shim_exec.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_call20.shim_exec.log
