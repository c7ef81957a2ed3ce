#!/bin/bash
# final check of the round: the whole GPU tier in collection order, then smoke()
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2_call23.pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_call23.pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_call23.smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_call23.smoke.log
