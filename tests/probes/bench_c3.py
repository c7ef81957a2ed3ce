"""BASELINE config 3: fit() objective for the rational-quadratic kernel, n = 5000, d = 4, 16 starts, sharded across
the ranks (torchrun) -- each rank runs the optimiser trajectories of its share of start vectors, no data-path
collective; the (par, value) pairs are all-gathered.  python tests/probes/bench_c3.py (1 GPU) or under torchrun."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ["GPRC_DEVICE"] = str(local)
import torch
import torch.distributed as dist

torch.cuda.set_device(local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD
import gprc_b200 as g
from oracle import gprc_oracle as o

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
cfg = o.make_config("C3", n=n)
ctx = g.Context(local)
obj = g.Objective(cfg["X"], cfg["y"], cfg["noise"], ctx=ctx, minors="cholesky")
obj.dens("rationalquadratic", [1.0, 1.0])  # warm-up
if world > 1:
    dist.barrier()
ctx.reset_timers()
t0 = time.perf_counter()
res = g.multistart(cfg["X"], cfg["y"], cfg["noise"], "rationalquadratic", cfg["starts"], ctx=ctx, group=group)
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
tm, launches = ctx.timers()
best = max(res, key=lambda r: r["value"])
if rank == 0:
    print(json.dumps(dict(config="C3 rationalquadratic n=%d d=4, 16 starts (BFGS, textbook gradient)" % n, world=world,
                          wall_s=wall, best_value=best["value"], best_par=list(best["par"]),
                          values=[round(r["value"], 4) for r in res], rank0_launches=launches)), flush=True)
if world > 1:
    dist.destroy_process_group()
