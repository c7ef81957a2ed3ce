#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shim_exec.py -x -q -m gpu > gpurun_out/r2_call19.pytest.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2_call19.pytest.log
