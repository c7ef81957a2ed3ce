#!/bin/bash
# ncu --set full of trsv_flow_kernel (alpha at n = 50 000), after the same command ran plain
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/round2/one_fit.py 50000 3 > gpurun_out/r2_call21.plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r2_call21.plain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trsv_flow_kernel -c 2 -f -o gpurun_out/r2_prof_trsv_flow \
  python tools/round2/one_fit.py 50000 1 > gpurun_out/r2_call21.ncu.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/r2_call21.ncu.log
ls -la gpurun_out | grep call21; ls -la gpurun_out/r2_prof_trsv_flow.ncu-rep
