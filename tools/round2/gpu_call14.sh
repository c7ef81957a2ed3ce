#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call14
python tools/potrf_probe.py > $O.plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:potrf_diag -s 40 -c 1 -o $O.prof_potrf_diag python tools/potrf_probe.py > $O.ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out | tail -2
