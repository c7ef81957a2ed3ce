"""torchrun --nproc-per-node N tools/dist_check.py [n ...]: distributed Cholesky + solve against the single-GPU path
(parity) and its throughput.  `--big n` skips the single-GPU comparison (matrix does not fit one device)."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["GPRC_DEVICE"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gprc_b200 as g
from importlib import import_module
DistGPR = import_module("gaussian-process-regression_b200.dist").DistGPR

ctx = g.Context(local)
D = DistGPR(ctx)
args = sys.argv[1:]
big = "--big" in args
kern = "sqrexp"
if "--kernel" in args:
    kern = args[args.index("--kernel") + 1]
sizes = [int(a) for a in args if a.isdigit()] or [3000, 8192]
specs = dict(sqrexp=(g.KernelSpec("sqrexp", l=1.0), 0.01), polynomial=(g.KernelSpec("polynomial", sigma=1.0, p=3.0), 0.1),
             gammaexp=(g.KernelSpec("gammaexp", l=1.0, gamma=1.5), 0.1))
spec, noise = specs[kern]
for n in sizes:
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (8, n))
    y = np.sum(np.sin(math.pi * X), axis=0) + rng.normal(0, 0.1, n)
    r = D.fit(X, y, noise, spec)          # warm-up (allocations, NCCL channels)
    dist.barrier()
    t0 = time.perf_counter()
    r = D.fit(X, y, noise, spec)
    dist.barrier()
    wall = time.perf_counter() - t0
    tf = n ** 3 / 3 / (r["phase_ms"]["factor"] * 1e-3) / 1e12
    out = dict(n=n, world=world, kernel=kern, info=r["info"], logp=r["logp"], phase_ms=r["phase_ms"], wall_s=wall,
               cholesky_tflops_aggregate=tf)
    if not big:
        kc, _ = spec.to_c()
        import ctypes as C
        h = C.c_void_p()
        lp, info = C.c_double(0.0), C.c_long(0)
        xp = np.ascontiguousarray(X.T)
        g._lib.check(ctx.lib.gprc_gpr_fit(ctx.handle, kc, g._lib.dptr(xp), 8, n, g._lib.dptr(y), noise, C.byref(h),
                                          C.byref(lp), C.byref(info)))
        a1 = np.empty(n)
        ctx.lib.gprc_gpr_get(h, 1, g._lib.dptr(a1))
        ctx.lib.gprc_gpr_free(h)
        out["logp_single"] = lp.value
        out["logp_rel_diff"] = abs(lp.value - r["logp"]) / abs(lp.value)
        out["alpha_max_rel_diff"] = float(np.max(np.abs(a1 - r["alpha"])) / np.max(np.abs(a1)))
    if rank == 0:
        print(json.dumps(out), flush=True)
D.close()
dist.destroy_process_group()
