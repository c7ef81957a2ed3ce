#!/bin/bash
# round 2, call 6 (1 GPU): dataflow TRSV v2 (tests + timing), S = 6 bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call6
timeout 300 python -m pytest tests/test_gpu_trsv_flow.py -m gpu -x -q > $O.pytest_trsv.log 2>&1; echo "pytest trsv rc=$?"; tail -5 $O.pytest_trsv.log
timeout 300 python tools/trsv_probe.py 2048 5000 16384 50000 > $O.trsv_probe.log 2>&1; echo "probe rc=$?"; cat $O.trsv_probe.log
for d in 0 16 64 80 2 18 66 82; do
  timeout 100 tools/oz_test time 7 16384 18944 63 5 0 $d 2>&1 | grep update_kernel
done > $O.oz_order.log; cat $O.oz_order.log
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --ozaki-digits 6 > $O.bench_c4_s6.json 2> $O.bench_c4_s6.err; echo "bench S=6 rc=$?"; tail -c 600 $O.bench_c4_s6.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench_c4_s6.json"))
    r=d["roofline"]
    print("C4 S=6 value %.3f e2e %.3f var_ms %.1f fp64eq %.1f" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"]))
    print("parity", d["parity"])
except Exception as e:
    print("no S=6 line", e)
PY
