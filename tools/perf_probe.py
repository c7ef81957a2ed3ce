"""Perf probe (not the bench): DGEMM core vs cuBLAS, blocked Cholesky vs cuSOLVER, phase times of a GPR fit+predict."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gprc_b200 as g
import torch

ctx = g.default_context()
lib = ctx.lib


def ev_time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.sync()
        t = time.perf_counter()
        fn()
        ctx.sync()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


what = sys.argv[1:] or ["gemm", "potrf", "gpr"]
if "gemm" in what:
    for n in (4096, 8192):
        A = torch.randn(n, n, dtype=torch.float64, device="cuda")
        B = torch.randn(n, n, dtype=torch.float64, device="cuda")
        Cc = torch.zeros(n, n, dtype=torch.float64, device="cuda")
        t_cublas = ev_time(lambda: torch.mm(A, B, out=Cc))
        for transb in (1, 0):
            t = ev_time(lambda: lib.gprc_dev_dgemm(ctx.handle, transb, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n,
                                                   0.0, Cc.data_ptr(), n))
            print("dgemm n=%d transb=%d: ours %.2f TF/s   cuBLAS %.2f TF/s" % (n, transb, 2 * n ** 3 / t / 1e12,
                                                                             2 * n ** 3 / t_cublas / 1e12), flush=True)
if "potrf" in what:
    for n in (8192, 16384, 32768):
        G = torch.randn(n, 256, dtype=torch.float64, device="cuda")
        A0 = G @ G.T / 256 + torch.eye(n, dtype=torch.float64, device="cuda") * 2
        A = A0.clone()
        dinv = torch.empty(n * 128, dtype=torch.float64, device="cuda")
        info = C.c_long(0)

        def ours():
            A.copy_(A0)
            torch.cuda.synchronize()
            ctx.reset_timers()
            lib.gprc_dev_potrf(ctx.handle, A.data_ptr(), n, n, dinv.data_ptr(), C.byref(info))
        ours()
        tm, _ = ctx.timers()
        ours()
        tm, launches = ctx.timers()
        t0 = time.perf_counter()
        torch.linalg.cholesky(A0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch.linalg.cholesky(A0)
        torch.cuda.synchronize()
        t_lib = time.perf_counter() - t0
        print("potrf n=%d: ours %.1f ms = %.2f TF/s (info %d, %d launches)  torch/cuSOLVER %.1f ms = %.2f TF/s" % (
            n, tm["chol"], n ** 3 / 3 / tm["chol"] / 1e9, info.value, launches, t_lib * 1e3, n ** 3 / 3 / t_lib / 1e12),
            flush=True)
        del A, A0, G
if "gpr" in what:
    n = int(sys.argv[sys.argv.index("gpr") + 1]) if len(sys.argv) > sys.argv.index("gpr") + 1 else 16384
    m = 65536
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (8, n))
    y = np.sum(np.sin(np.pi * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, (8, m))
    ctx.reset_timers()
    t0 = time.perf_counter()
    model = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0))
    t1 = time.perf_counter()
    pred = model.predict(Xs)
    t2 = time.perf_counter()
    tm, launches = ctx.timers()
    print("gpr n=%d m=%d: fit %.3f s predict %.3f s launches %d timers %s" % (n, m, t1 - t0, t2 - t1, launches,
                                                                           {k: round(v, 2) for k, v in tm.items()}))
    print("  chol %.2f TF/s  trtri %.2f TF/s  var %.2f TF/s" % (n ** 3 / 3 / tm["chol"] / 1e9, (n ** 3 / 3 / tm["trtri"] / 1e9) if tm["trtri"] else float("nan"),
                                                                 n * n * m / tm["var"] / 1e9))
