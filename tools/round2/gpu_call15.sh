#!/bin/bash
# round 2, call 15 (4 GPUs): the driver's invocation on 4 and on 2 of the box's GPUs with the final defaults
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call15
for N in 4 2; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > $O.bench_${N}gpu.json 2> $O.bench_${N}gpu.err ) 2> $O.time$N; echo "bench$N rc=$?"; tail -3 $O.time$N | head -1
python - <<PY
import json
try:
    line=[l for l in open("$O.bench_${N}gpu.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line)
    r=d["roofline"]
    print("$N GPUs: value %.3f e2e %.3f var_ms %.1f fp64eq/GPU %.1f chol %.1f TF agg; chunks: %s" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], d["cholesky_tflops"], r["algorithmic_bytes_note"][-22:]))
    print("  parity", d["parity"]["max_rel_dvar_vs_max_var_kss"], d["parity"]["max_rel_dmean"], "clocks", d["clocks"])
except Exception as e:
    print("no line", e)
PY
done
