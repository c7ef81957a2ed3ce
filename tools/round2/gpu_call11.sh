#!/bin/bash
# round 2, call 11 (8 GPUs): the driver's 8-GPU invocation with the final defaults
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call11
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > $O.bench_8gpu.json 2> $O.bench_8gpu.err ) 2> $O.time; echo "bench8 rc=$?"; tail -3 $O.time; tail -c 300 $O.bench_8gpu.err
python - <<PY
import json
try:
    line=[l for l in open("$O.bench_8gpu.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line)
    r=d["roofline"]
    print("8 GPUs: value %.3f e2e %.3f var_ms %.1f fp64eq/GPU %.1f chol %.1f TF agg" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], d["cholesky_tflops"]))
    print("phases", d["phase_ms_per_step"]); print("parity", d["parity"]); print("clocks", d["clocks"]); print(r["algorithmic_bytes_note"])
except Exception as e:
    print("no 8-GPU line", e)
PY
