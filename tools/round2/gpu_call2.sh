#!/bin/bash
# round 2, second GPU call: whole GPU tier (golden at full size included), full-size C4 bench, launch list + full capture
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call2
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O.pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $O.pytest.log
GPRC_TRACE_CHUNKS=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --train-size 16384 --test-size 303104 --int8-tile 1 \
    > $O.bench_n16k_v1.json 2> $O.bench_n16k_v1.err; echo "bench n16k v1 rc=$?"; tail -12 $O.bench_n16k_v1.err
timeout 900 python bench.py --steps 3 --warmup 1 > $O.bench_c4.json 2> $O.bench_c4.err; echo "bench c4 rc=$?"; tail -c 600 $O.bench_c4.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench_c4.json"))
    r=d["roofline"]
    print("C4 value %.3f e2e %.3f var_ms %.1f fp64eq %.1f int8 %.0f of %.0f (frac %.2f)" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], r["achieved"], r["peak"], r["frac"]))
    print("clocks", d["clocks"]); print("parity", d["parity"]); print("phases", d["phase_ms_per_step"]); print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no C4 line", e)
PY
# launch list of a reduced step (the full C4 step has ~20k launches) -- plain run first, then under ncu
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-sample 0 --int8-peak-seconds 0.05 --train-size 16384 --test-size 151552"
$CMD > $O.plain_n16k.json 2> $O.plain_n16k.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O.launches_n16k.csv $CMD > $O.ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l $O.launches_n16k.csv
tools/oz_test time 7 16384 18944 63 5 0 > $O.oz_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:update_stack -s 100 -c 1 -o $O.prof_stack_pair tools/oz_test time 7 16384 18944 63 5 0 > $O.ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/ | tail -5
