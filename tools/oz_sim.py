"""NumPy emulation of the INT8 (Ozaki) blocked substitution of csrc/ozaki.cuh: the same fixed-point digits (bias trick),
the same order truncation (pairs with a + b >= S dropped), per-row exponents for L from the row maximum, per-test-point
exponents for V from sqrt(k**), Horner combination in FP64.  Digit products are accumulated in float64 matmuls, which
is exact here (|sum| < 2^53).  Used to choose S before the kernel existed and as a CPU-tier test of the scheme's
accuracy (tests/test_oracle.py).

    python tools/oz_sim.py 1024 7
"""
import sys

import numpy as np
import scipy.linalg as sl

NB = 128


def digits(x, e, S):
    """x = 2^(e - 8S + 2) sum_t d_t 256^(S-1-t), d_t in [-128, 127]; returns [d_0 (top), ..., d_{S-1}] as float64."""
    X = np.rint(np.ldexp(x, (8 * S - 2) - e)).astype(np.int64)
    bias = sum(0x80 << (8 * t) for t in range(S - 1))
    Y = X + bias
    out = []
    for s in range(S):
        t = S - 1 - s
        if t == S - 1:
            d = Y >> (8 * t)
        else:
            d = ((Y >> (8 * t)) & 0xFF) - 128
        out.append(d.astype(np.float64))
    rec = np.zeros_like(X)
    for d in out:
        rec = rec * 256 + d.astype(np.int64)
    assert np.array_equal(rec, X) and min(d.min() for d in out) >= -128 and max(d.max() for d in out) <= 127
    return out


def exponent_of(mx):
    e = np.frexp(mx)[1]
    return np.where(mx > 0, e, 0)


def substitution_int8(L, Ks, kss, S):
    """V = L^-1 Ks by blocked substitution with the off-diagonal products in digit arithmetic; returns V."""
    n, m = Ks.shape
    assert n % NB == 0
    nb = n // NB
    eL = np.zeros(n, dtype=np.int64)
    for i in range(1, nb):
        eL[i * NB:(i + 1) * NB] = exponent_of(np.abs(L[i * NB:(i + 1) * NB, :i * NB]).max(axis=1))
    eV = exponent_of(np.sqrt(np.maximum(kss, 0.0)) * (1.0 + 1e-9))
    Ld = [np.zeros((n, n)) for _ in range(S)]
    for i in range(1, nb):
        r = slice(i * NB, (i + 1) * NB)
        for s, d in enumerate(digits(L[r, :i * NB], eL[r, None], S)):
            Ld[s][r, :i * NB] = d
    T = Ks.copy()
    V = np.zeros_like(Ks)
    Vd = [np.zeros_like(Ks) for _ in range(S)]
    for i in range(nb):
        r = slice(i * NB, (i + 1) * NB)
        if i > 0:
            acc = [np.zeros((NB, m)) for _ in range(S)]
            for a in range(S):
                for b in range(S - a):
                    acc[a + b] += Ld[a][r, :i * NB] @ Vd[b][:i * NB]
            h = acc[S - 1].copy()
            for o in range(S - 2, -1, -1):
                h = h * 2.0 ** -8 + acc[o]
            T[r] -= np.ldexp(1.0, eL[r] - 6)[:, None] * np.ldexp(1.0, eV - 6)[None, :] * h
        V[r] = sl.solve_triangular(L[r, r], T[r], lower=True)
        assert np.all(np.abs(np.ldexp(V[r], -eV[None, :])) < 1.9), "V left its fixed-point range"
        for s, d in enumerate(digits(V[r], eV[None, :], S)):
            Vd[s][r] = d
    return V


def experiment(n, m, S, d=8, seed=4, noise=0.01):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (d, n))
    Xs = rng.uniform(-1, 1, (d, m))
    k = lambda A, B: np.exp(-((A[:, :, None] - B[:, None, :]) ** 2).sum(0) / 2)
    L = sl.cholesky(k(X, X) + noise * np.eye(n), lower=True)
    Ks = k(X, Xs)
    V64 = sl.solve_triangular(L, Ks, lower=True)
    V8 = substitution_int8(L, Ks, np.ones(m), S)
    var64, var8 = 1 - (V64 ** 2).sum(0), 1 - (V8 ** 2).sum(0)
    return np.abs(V8 - V64).max(), np.abs(var8 - var64).max(), var64.min()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    dv, dvar, vmin = experiment(n, 256, S)
    print("n %d S %d: max|V - V_fp64| %.3e  max|var - var_fp64| %.3e  (min var %.3e)" % (n, S, dv, dvar, vmin))
