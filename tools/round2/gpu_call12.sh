#!/bin/bash
# round 2, call 12 (1 GPU): full ncu capture of the final stacked-pair kernel (widest MMAs last)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call12
tools/oz_test time 7 16384 18944 63 5 0 > $O.oz_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:update_stack -s 100 -c 1 -o $O.prof_stack_pair_final tools/oz_test time 7 16384 18944 63 5 0 > $O.ncu_full.log 2>&1
echo "full capture rc=$?"; cat $O.oz_plain.log | tail -2; ls -la gpurun_out | tail -3
