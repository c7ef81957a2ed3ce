#!/bin/bash
# round 2, first GPU call: stacked-plane INT8 kernels (checks against the digit emulation, rates), the whole GPU tier,
# reduced-size bench per kernel variant
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call1
: > $O.oz.log
for a in "check 7 1024 256 3 4 0" "check 7 2048 256 9 4 0" "check 6 1024 128 5 4 0" "check 8 1024 128 5 4 0" \
         "check 7 1024 256 1 5 0" "check 7 1024 256 3 5 0" "check 7 2048 384 7 5 0" "check 8 16640 128 64 5 0" "check 7 16896 128 131 4 0" \
         "time 7 16384 18944 127 0 0" "time 7 16384 18944 127 4 0" "time 7 16384 18944 63 5 0" "time 7 16384 18944 127 2 0" \
         "time 7 16384 18944 127 4 0 1" "time 7 16384 18944 127 4 0 2" "time 7 16384 18944 63 5 0 1" "time 7 16384 18944 63 5 0 2" \
         "time 6 16384 18944 127 4 0" "time 6 16384 18944 63 5 0" "time 8 16384 18944 63 5 0"; do
  echo "== oz_test $a" >> $O.oz.log
  timeout 300 tools/oz_test $a 2>&1 | grep -E "RESULT|update_kernel|mismatch|error|failed|stacked" >> $O.oz.log
done
tail -40 $O.oz.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O.pytest.log
for v in 64 1 2; do
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --train-size 16384 --test-size 303104 --int8-tile $v \
    > $O.bench_n16k_v$v.json 2> $O.bench_n16k_v$v.err; echo "bench n16k v$v rc=$?"; tail -c 300 $O.bench_n16k_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O.bench_n16k_v$v.json"))
    r=d["roofline"]
    print("v$v value %.3f e2e %.3f var_ms %.1f fp64eq %.1f int8 %.0f of %.0f (frac %.2f) clocks %s parity %s" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], r["achieved"], r["peak"], r["frac"], d["clocks"], d["parity"]))
except Exception as e:
    print("v$v: no line", e)
PY
done
