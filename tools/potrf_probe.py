"""one GPR fit at n = 2048 through the launch-sequence Cholesky (for an ncu capture of potrf_diag_kernel)"""
import sys
import numpy as np
sys.path.insert(0, ".")
import gprc_b200 as g
ctx = g.default_context()
ctx.set_option(g._lib.OPT_CHOL_TILES, 0)
rng = np.random.default_rng(1)
X = rng.uniform(-1, 1, (8, 2048))
y = np.sum(np.sin(3 * X), axis=0)
for _ in range(3):
    m = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0), ctx=ctx)
print(m.logp)
