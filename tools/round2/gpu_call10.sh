#!/bin/bash
# round 2, call 10 (1 GPU): the driver's sequence -- GPU tier, smoke, bench with --steps 20 --warmup 5
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call10
( time timeout 1100 python -m pytest tests -x -q -m gpu > $O.pytest.log 2>&1 ) 2> $O.pytest.time; echo "pytest rc=$?"; tail -3 $O.pytest.log; tail -3 $O.pytest.time
timeout 300 python __graft_entry__.py smoke > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O.smoke.log
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O.bench.json 2> $O.bench.err ) 2> $O.bench.time; echo "bench rc=$?"; cat $O.bench.time | tail -3; tail -c 400 $O.bench.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench.json"))
    r=d["roofline"]
    print("C4 value %.3f e2e %.3f var_ms %.1f fp64eq %.1f int8 %.0f of %.0f (frac %.2f) chol %.1f TF" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], r["achieved"], r["peak"], r["frac"], d["cholesky_tflops"]))
    print("clocks", d["clocks"]); print("parity", d["parity"]); print("phases", d["phase_ms_per_step"]); print(r["algorithmic_bytes_note"], r["traffic"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["literal_value"])
except Exception as e:
    print("no line", e)
PY
