#!/bin/bash
# round 2, call 8 (1 GPU): persistent tile Cholesky (tests + timing)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call8
timeout 600 python -m pytest tests/test_gpu_chol_persistent.py -m gpu -x -q > $O.pytest_chol.log 2>&1; echo "pytest chol rc=$?"; tail -15 $O.pytest_chol.log
timeout 400 python tests/probes/chol_probe.py 2048 5000 8192 16384 24576 > $O.chol_probe.log 2>&1; echo "probe rc=$?"; cat $O.chol_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O.pytest.log
timeout 300 python tests/probes/bench_small.py C1 C2 C3 > $O.small.log 2>&1; cat $O.small.log
