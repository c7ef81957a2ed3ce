"""Full-size golden sample for BASELINE config 4 (n = 50 000, d = 8, sqrexp l = 1, noise 0.01): the ORACLE's arithmetic
at full size on the first test points of bench.py's generator, so that the timed GPU result is pinned to the reference's
restatement instead of to the library's own FP64 path.

    python tests/golden/make_c4_golden.py            # ~15-25 min on 8 cores, 21 GB of RAM; writes c4_full_sample.npz

What it computes is R/GPRclass.R:138-164 literally, memory-bounded (the reference's outer() gather cannot exist at
this size):
  K      = covariance_matrix(X, X, k) + noise * diag(n)   column blocks through oracle.sqrexp on gathered columns
           (the same colSums((x - y)^2) / exp order as the oracle, only chunked), lower triangle + diagonal
  L      = t(chol(K))                                      oracle.blocked_cholesky_inplace: dpotrf's blocked algorithm over
           LAPACK / BLAS calls on 2048-blocks (SciPy's LP64 LAPACK segfaults for n^2 > 2^31), in place
  alpha  = solve(t(L), solve(L, y))                        block substitution over dtrtrs / dgemm ("generous": triangular,
           SURVEY.md A.7)
  logp   = -0.5 y'alpha - sum(log(diag(L))) - n/2 log(2 pi)
  K_star, mean = t(K_star) alpha, v = solve(L, K_star), var = k(X*, X*) - colSums(v * v)   for the first M test points
Stored: mean, var, logp, alpha[:64], diag(L)[:64] and a few sampled entries of L, plus the generator's parameters.
PARITY UNPINNED by the reference itself (the oracle's header explains why); pinned by this restatement.
"""
import math
import os
import sys
import time

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gprc_oracle as o  # noqa: E402


def make_inputs(n, m, d):  # bench.py:make_inputs, bit for bit (seed 4; the test points are drawn after X and y)
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, size=(d, n))
    y = np.sum(np.sin(math.pi * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, size=(d, m))
    return X, y, Xs


def main(n=50000, m=1000000, d=8, M=256, noise=0.01, l=1.0, out=None, block=250):
    t0 = time.time()
    X, y, Xs = make_inputs(n, m, d)
    Xs = np.ascontiguousarray(Xs[:, :M])
    K = np.zeros((n, n), order="F")
    for j0 in range(0, n, block):                     # column block j0 .. j1 of the lower triangle
        j1 = min(n, j0 + block)
        rows = n - j0
        # entry (i, j) = k(X[, i], X[, j]): the two gathered operands of outer() (R/GPRclass.R:356), one column block at
        # a time, through the oracle's .matrix kernel
        xa = np.broadcast_to(X[:, j0:, None], (d, rows, j1 - j0)).reshape(d, -1)
        xb = np.broadcast_to(X[:, None, j0:j1], (d, rows, j1 - j0)).reshape(d, -1)
        K[j0:, j0:j1] = o.sqrexp(xa, xb, l).reshape(rows, j1 - j0)
        if (j0 // block) % 20 == 0:
            print("K build: column %d of %d, %.0f s" % (j0, n, time.time() - t0), flush=True)
    K[np.arange(n), np.arange(n)] += noise            # K + noise * diag(n)                 R/GPRclass.R:142
    print("K built, %.0f s" % (time.time() - t0), flush=True)
    # scipy.linalg.cholesky (LP64 LAPACK in this image) segfaults for n^2 > 2^31: the oracle's blocked dpotrf instead
    L = o.blocked_cholesky_inplace(K)
    print("dpotrf (blocked, in place) done, %.0f s" % (time.time() - t0), flush=True)
    # its strict upper triangle is garbage (the untouched upper part of K is zero here: only the lower triangle was built)
    z = o.blocked_solve_lower(L, y)
    alpha = o.blocked_solve_lower(L, z, trans=True)                                           # :152
    dg = np.diag(L).copy()
    logp = -0.5 * float(y @ alpha) - float(np.sum(np.log(dg))) - n / 2 * math.log(2 * math.pi)  # :153
    k = o.cov_func(o.sqrexp, l=l)
    Ks = o.covariance_matrix(X, Xs, k)                                                         # :160
    mean = Ks.T @ alpha                                                                        # :161
    v = o.blocked_solve_lower(L, Ks)                                                           # :162
    var = k(Xs, Xs) - np.sum(v * v, axis=0)                                                    # :164
    rng = np.random.default_rng(99)
    ii = rng.integers(0, n, 64)
    jj = np.minimum(ii, rng.integers(0, n, 64))
    out = out or os.path.join(HERE, "c4_full_sample.npz")
    np.savez(out, n=n, m=m, d=d, M=M, noise=noise, l=l, seed=4, mean=mean, var=var, logp=logp, alpha_head=alpha[:64],
             diagL_head=dg[:64], L_rows=ii, L_cols=jj, L_vals=L[ii, jj], Xs_head=Xs,
             yalpha=float(y @ alpha), sumlogdiag=float(np.sum(np.log(dg))))
    print("wrote %s: logp = %.12f, mean[0] = %.12f, var[0] = %.12e, total %.0f s"
          % (out, logp, mean[0], var[0], time.time() - t0), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:   # small self-check: python make_c4_golden.py <n> compares with oracle.GPR at that size
        n = int(sys.argv[1])
        main(n=n, m=1000, M=64, out="/tmp/c4_small.npz", block=100)
        X, y, Xs = make_inputs(n, 1000, 8)
        g = o.GPR(X, y, 0.01, o.cov_func(o.sqrexp, l=1.0))
        p = g.predict(Xs[:, :64])
        z = np.load("/tmp/c4_small.npz")
        print("self-check vs oracle.GPR: dmean %.2e dvar %.2e dlogp %.2e" % (np.max(np.abs(p[:, 0] - z["mean"])),
                                                                             np.max(np.abs(p[:, 1] - z["var"])),
                                                                             abs(float(g.logp) - float(z["logp"]))))
    else:
        main()
