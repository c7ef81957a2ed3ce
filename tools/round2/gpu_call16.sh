#!/bin/bash
# round 2, last GPU call: the GPU tier, smoke, and bench.py with no flags
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call16
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=5 > $O.pytest.log 2>&1 ) 2> $O.pytest.time; echo "pytest rc=$?"; tail -9 $O.pytest.log; tail -3 $O.pytest.time | head -1
timeout 200 python __graft_entry__.py smoke > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O.smoke.log | head -c 200; echo
( time timeout 400 python bench.py > $O.bench_default.json 2> $O.bench_default.err ) 2> $O.bench.time; echo "bench rc=$?"; tail -3 $O.bench.time | head -1; tail -c 300 $O.bench_default.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench_default.json"))
    print("default bench: value %.3f e2e %.3f steps %d warmup %d; cpu_baseline %.0f s (literal %.0f) on %d cores" % (d["value"], d["e2e"]["value"], d["steps"], d["warmup"], d["cpu_baseline"]["value"], d["cpu_baseline"]["literal_value"], d["cpu_baseline"]["cores"]))
    print(d["cpu_baseline"]["sample"][:400])
except Exception as e:
    print("no line", e)
PY
