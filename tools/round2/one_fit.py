"""One GPR fit (K build, Cholesky, alpha + logp) at the given n with the default options: the command the round-2 ncu
captures of the dataflow substitution kernel and of the persistent tile Cholesky wrap."""
import sys

import numpy as np

sys.path.insert(0, ".")
import gprc_b200 as g

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = g.default_context()
rng = np.random.default_rng(n)
X = rng.uniform(-1, 1, (8, n))
y = np.sum(np.sin(3 * X), axis=0) + rng.normal(0, 0.1, n)
for rep in range(reps):
    ctx.reset_timers()
    m = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0), ctx=ctx)
    t, _ = ctx.timers()
    print("n=%d rep %d: build %.3f ms, chol %.3f ms, solve %.3f ms, logp %.9g" % (n, rep, t["build_k"], t["chol"], t["solve"],
                                                                               float(np.asarray(m.logp).ravel()[0])), flush=True)
    del m
