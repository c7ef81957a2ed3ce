#!/bin/bash
# ncu --set full of chol_persistent_kernel (whole Cholesky at n = 8192 as one launch), after the same command ran plain
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/round2/one_fit.py 8192 3 > gpurun_out/r2_call22.plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r2_call22.plain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chol_persistent_kernel -c 1 -f -o gpurun_out/r2_prof_chol_persistent \
  python tools/round2/one_fit.py 8192 1 > gpurun_out/r2_call22.ncu.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/r2_call22.ncu.log
ls -la gpurun_out/r2_prof_chol_persistent.ncu-rep
