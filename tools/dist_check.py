"""torchrun --nproc-per-node N tools/dist_check.py [n ...]: distributed Cholesky + solve against the single-GPU path
(parity) and its throughput.  `--big n` skips the single-GPU comparison (matrix does not fit one device).
`--timeline FILE` records the per-panel timeline of the timed factorisation on every rank (gprc_dist_set_timeline) and
writes, on rank 0, the raw events of all ranks plus a summary: how long the main stream of each rank spent in trailing
updates, how long it waited for a panel (exposed factorisation + broadcast time -- what a 2-D layout could shorten), and
the bytes every rank received."""
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
timeline_file = args[args.index("--timeline") + 1] if "--timeline" in args else None
out_file = args[args.index("--out") + 1] if "--out" in args else None
sizes = [int(a) for a in args if a.isdigit()] or [3000, 8192]
records = []
specs = dict(sqrexp=(g.KernelSpec("sqrexp", l=1.0), 0.01), polynomial=(g.KernelSpec("polynomial", sigma=1.0, p=3.0), 0.1),
             gammaexp=(g.KernelSpec("gammaexp", l=1.0, gamma=1.5), 0.1))
spec, noise = specs[kern]
for n in sizes:
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (8, n))
    y = np.sum(np.sin(math.pi * X), axis=0) + rng.normal(0, 0.1, n)
    r = D.fit(X, y, noise, spec)          # warm-up (allocations, NCCL channels)
    dist.barrier()
    if timeline_file:
        g._lib.check(ctx.lib.gprc_dist_set_timeline(D.handle, 1))
    t0 = time.perf_counter()
    r = D.fit(X, y, noise, spec)
    dist.barrier()
    wall = time.perf_counter() - t0
    tl = None
    if timeline_file:
        import ctypes as C
        npan = C.c_int(0)
        g._lib.check(ctx.lib.gprc_dist_get_timeline(D.handle, None, 0, C.byref(npan)))
        mine = np.full((npan.value, 6), np.nan)
        g._lib.check(ctx.lib.gprc_dist_get_timeline(D.handle, g._lib.dptr(mine), npan.value, C.byref(npan)))
        g._lib.check(ctx.lib.gprc_dist_set_timeline(D.handle, 0))
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank == 0:
            T = np.stack(gathered)                              # world x npan x 6
            fac = r["phase_ms"]["factor"]
            upd = np.nansum(T[:, :, 5] - T[:, :, 4], axis=1)    # ms each rank's main stream spent in unpack + trailing update
            own = np.array([np.nansum((T[w, :, 1] - T[w, :, 0])) for w in range(world)])   # look-ahead + panel factorisation
            bc = np.nanmax(T[:, :, 3] - T[:, :, 2], axis=0)     # per panel: slowest rank's broadcast
            n_pad = (n + 511) // 512 * 512
            rows = n_pad - 512 * np.arange(npan.value)
            recv = float(np.sum(rows * 512 * 8))
            tl = dict(npan=int(npan.value), factor_ms=fac,
                      main_stream_update_ms_per_rank=[float(v) for v in upd],
                      main_stream_waiting_ms_per_rank=[float(fac - v) for v in upd],
                      main_stream_waiting_share_max=float(np.max(fac - upd) / fac),
                      owner_lookahead_plus_factor_ms_per_rank=[float(v) for v in own],
                      broadcast_ms_sum_over_panels=float(np.nansum(bc)),
                      broadcast_ms_first_panels=[float(v) for v in bc[:4]], broadcast_ms_last_panels=[float(v) for v in bc[-4:]],
                      bytes_received_per_rank=recv,
                      broadcast_gbs_effective=float(recv / (np.nansum(bc) * 1e-3) / 1e9),
                      note="waiting = factorisation time not covered by unpack + trailing updates on that rank's main stream: "
                           "the exposed part of panel factorisation + broadcast, i.e. the upper bound of what a 2-D "
                           "block-cyclic layout (shorter panels per owner) could recover")
            with open(timeline_file.replace(".json", "_%s_n%d.json" % (kern, n)), "w") as f:
                json.dump(dict(n=n, world=world, kernel=kern, summary=tl, columns=["la_factor_begin", "la_factor_end",
                               "bcast_begin", "bcast_end", "update_begin", "update_end"],
                               events_ms=np.where(np.isnan(T), -1.0, np.round(T, 3)).tolist()), f)
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
    if tl:
        out["timeline"] = tl
    if rank == 0:
        print(json.dumps(out), flush=True)
        records.append(out)
if rank == 0 and out_file:
    with open(out_file, "w") as f:
        json.dump(records, f, indent=1)
D.close()
dist.destroy_process_group()
