#!/bin/bash
# round 2, call 7 (1 GPU): reversed MMA order as default (bit-exact checks), TRSV v2.1 default, whole GPU tier, C4 bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call7
for a in "check 7 1024 256 3 4 0" "check 7 2048 384 7 5 0" "check 8 16640 128 64 5 0" "check 6 1024 256 3 5 0" "check 7 16896 128 131 4 0" "time 7 16384 18944 63 5 0" "time 7 16384 18944 127 4 0"; do
  echo "== oz_test $a"; timeout 300 tools/oz_test $a 2>&1 | grep -E "RESULT|update_kernel|mismatch|error|failed"
done > $O.oz.log 2>&1; cat $O.oz.log
timeout 300 python tools/trsv_probe.py 2048 5000 16384 50000 > $O.trsv_probe.log 2>&1; echo "probe rc=$?"; cat $O.trsv_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O.pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O.bench_c4.json 2> $O.bench_c4.err; echo "bench c4 rc=$?"; tail -c 600 $O.bench_c4.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench_c4.json"))
    r=d["roofline"]
    print("C4 value %.3f e2e %.3f var_ms %.1f fp64eq %.1f int8 %.0f of %.0f (frac %.2f)" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], r["achieved"], r["peak"], r["frac"]))
    print("clocks", d["clocks"]); print("parity", d["parity"]); print("phases", d["phase_ms_per_step"])
except Exception as e:
    print("no C4 line", e)
PY
