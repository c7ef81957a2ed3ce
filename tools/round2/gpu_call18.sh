#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_primitives.py -x -q -m gpu > gpurun_out/r2_call18.pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2_call18.pytest.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1 | head -c 150
