"""torchrun --nproc-per-node 2 tools/dist_predict_check.py: distributed train + replicated factor, then predict;
compared with the single-GPU path on the same inputs."""
import ctypes as C
import math
import os
import sys

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
for n, m in [(1000, 300), (5000, 2000), (16384, 4096)]:
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, (8, n))
    y = np.sum(np.sin(math.pi * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, (8, m))
    spec = g.KernelSpec("sqrexp", l=1.0)
    xp, xsp = np.ascontiguousarray(X.T), np.ascontiguousarray(Xs.T)
    h, lp, info, ph = D.fit_replicated(xp, 8, n, y, 0.01, spec)
    mean, var = np.empty(m), np.empty(m)
    g._lib.check(ctx.lib.gprc_gpr_predict(h, g._lib.dptr(xsp), m, g._lib.dptr(mean), g._lib.dptr(var)))
    ctx.lib.gprc_gpr_free(h)
    # like with like: the distributed factorisation follows the launch-sequence algorithm (512-column panels), so the
    # single-GPU reference is factored by that one too -- the persistent tile kernel that single-GPU fits of this size
    # take by default sums every tile in one pass (same result to 1e-12, not the same bits)
    ctx.set_option(g._lib.OPT_CHOL_TILES, 0)
    try:
        ref = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0), ctx=ctx)
    finally:
        ctx.set_option(g._lib.OPT_CHOL_TILES, g._lib.CHOL_TILES_DEFAULT)
    pr = ref.predict(Xs)
    dflt = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0), ctx=ctx).predict(Xs)
    assert np.max(np.abs(dflt[:, 0] - pr[:, 0])) <= 1e-9 * np.max(np.abs(pr[:, 0])) and np.max(np.abs(dflt[:, 1] - pr[:, 1])) <= 1e-9
    print("rank %d n=%d: info %d logp rel diff %.2e  max|dmean| %.2e  max|dvar| %.2e  factor %.1f ms" % (
        rank, n, info, abs(lp - ref.logp[0, 0]) / abs(lp), np.max(np.abs(mean - pr[:, 0])), np.max(np.abs(var - pr[:, 1])),
        ph["factor"]), flush=True)
D.close()
dist.destroy_process_group()
