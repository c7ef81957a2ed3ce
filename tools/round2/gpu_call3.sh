#!/bin/bash
# round 2, third GPU call (1 GPU): the driver's own invocations -- reference arm, then the bench with --steps 20 --warmup 5
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call3
nproc; free -g | head -2
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O.reference.json 2> $O.reference.err ) 2> $O.reference.time; echo "reference rc=$?"; cat $O.reference.time | tail -3; tail -c 400 $O.reference.err; head -c 1800 $O.reference.json; echo
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O.bench.json 2> $O.bench.err ) 2> $O.bench.time; echo "bench rc=$?"; cat $O.bench.time | tail -3; tail -c 400 $O.bench.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench.json"))
    r=d["roofline"]
    print("C4 value %.3f e2e %.3f var_ms %.1f fp64eq %.1f int8 %.0f of %.0f (frac %.2f) chol %.1f TF" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], r["achieved"], r["peak"], r["frac"], d["cholesky_tflops"]))
    print("clocks", d["clocks"]); print("parity", d["parity"]); print("phases", d["phase_ms_per_step"]); print(r["algorithmic_bytes_note"], r["peaks"]["measured_sustained"], r["traffic"])
except Exception as e:
    print("no line", e)
PY
