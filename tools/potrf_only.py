import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
import gprc_b200 as g
ctx = g.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
G = rng.standard_normal((n, 64))
A = G @ G.T / 64 + 2 * np.eye(n)
dA = ctx.upload(np.asfortranarray(A))
dinv = ctx.malloc(n * 128 * 8)
info = C.c_long(0)
for _ in range(2):
    ctx.h2d(dA, np.asfortranarray(A))
    ctx.reset_timers()
    ctx.lib.gprc_dev_potrf(ctx.handle, dA, n, n, dinv, C.byref(info))
    print(ctx.timers())
