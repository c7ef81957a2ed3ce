#!/bin/bash
# round 2, 8-GPU call: dist tests, C4 bench on 8 GPUs, C5 (n = 200 000) factorisation with the per-panel timeline
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out/r2_call4
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > $O.pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 $O.pytest_dist.log
timeout 600 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > $O.bench_8gpu.json 2> $O.bench_8gpu.err; echo "bench8 rc=$?"; tail -c 500 $O.bench_8gpu.err
python - <<PY
import json
try:
    d=json.load(open("$O.bench_8gpu.json"))
    r=d["roofline"]
    print("8 GPUs: value %.3f e2e %.3f var_ms %.1f fp64eq/GPU %.1f chol %.1f TF agg; phases %s; parity %s" % (d["value"], d["e2e"]["value"], r["ms_per_step"], r["fp64_equivalent"]["achieved"], d["cholesky_tflops"], d["phase_ms_per_step"], d["parity"]))
except Exception as e:
    print("no 8-GPU line", e)
PY
for k in polynomial gammaexp; do
  timeout 600 $TR --nproc-per-node 8 --master-port 29513 tools/dist_check.py --big 200000 --kernel $k --timeline $O.timeline.json --out $O.dist_c5_$k.json > $O.dist_c5_$k.log 2>&1; echo "dist_check $k rc=$?"; tail -c 1500 $O.dist_c5_$k.log
done
timeout 300 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 2 --warmup 3 > $O.bench_4gpu.json 2> $O.bench_4gpu.err; echo "bench4 rc=$?"; head -c 300 $O.bench_4gpu.json
ls -la gpurun_out | tail -8
