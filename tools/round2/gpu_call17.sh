#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_gpc_fit.py -x -q -m gpu > gpurun_out/r2_call17.pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_call17.pytest.log
