import numpy as np, scipy.linalg as sl, sys
rng = np.random.default_rng(4)
n, m, d = int(sys.argv[1]), 256, 8
S = int(sys.argv[2])
X = rng.uniform(-1, 1, (d, n)); Xs = rng.uniform(-1, 1, (d, m))
def k(A, B): 
    r2 = ((A[:, :, None] - B[:, None, :])**2).sum(0); return np.exp(-r2 / 2)
K = k(X, X) + 0.01 * np.eye(n); Ks = k(X, Xs)
L = sl.cholesky(K, lower=True)
V = sl.solve_triangular(L, Ks, lower=True)
var_ref = 1 - (V**2).sum(0)
# long double reference of the substitution
def slices(x, e, S):
    B = 8 * S
    Xi = np.rint(np.ldexp(x, (B - 2) - e)).astype(object)  # python ints
    out = []
    bias = sum(0x80 << (8 * t) for t in range(S - 1))
    Y = Xi + bias
    for t in range(S):  # t = 0 lowest byte
        if t < S - 1:
            b = np.array([(int(v) >> (8 * t)) & 0xFF for v in Y.ravel()], dtype=np.int64).reshape(x.shape) - 128
        else:
            b = np.array([int(v) >> (8 * t) for v in Y.ravel()], dtype=np.int64).reshape(x.shape)
        out.append(b.astype(np.float64))
    return out[::-1]  # top first
NBk = 128
nb = n // NBk
eL = np.zeros(n, dtype=int)
for r in range(n):
    mx = np.abs(L[r, :(r // NBk) * NBk]).max() if r >= NBk else 0.0
    eL[r] = np.frexp(mx)[1] if mx > 0 else 0
eV = np.frexp(np.sqrt(np.ones(m)) * 1.0001)[1]
Ls = slices(L, eL[:, None], S)
# check reconstruction
rec = sum(Ls[t] * 2.0**(8 * (S - 1 - t)) for t in range(S)) * np.ldexp(1.0, eL[:, None] - (8 * S - 2))
print("L recon err", np.abs(np.tril(rec - L, -NBk)).max(), "digit range", min(a.min() for a in Ls), max(a.max() for a in Ls))
T = Ks.copy()
Vz = np.zeros_like(Ks)
Vs = [np.zeros_like(Ks) for _ in range(S)]
for i in range(nb):
    r0, r1 = i * NBk, (i + 1) * NBk
    if i > 0:
        acc = [np.zeros((NBk, m)) for _ in range(S)]
        for a in range(S):
            for b in range(S - a):
                acc[a + b] += Ls[a][r0:r1, :r0] @ Vs[b][:r0, :]
        h = acc[S - 1].copy()
        for o in range(S - 2, -1, -1):
            h = h * 2.0**-8 + acc[o]
        sc_r = np.ldexp(1.0, 8 * (S - 1) - ((8 * S - 2) - eL[r0:r1]))
        sc_c = np.ldexp(1.0, 8 * (S - 1) - ((8 * S - 2) - eV))
        T[r0:r1] -= sc_r[:, None] * sc_c[None, :] * h
    Vi = sl.solve_triangular(L[r0:r1, r0:r1], T[r0:r1], lower=True)
    Vz[r0:r1] = Vi
    sl_ = slices(Vi, eV[None, :], S)
    for t in range(S): Vs[t][r0:r1] = sl_[t]
var_oz = 1 - (Vz**2).sum(0)
print("n", n, "S", S, "max|V-Vref|", np.abs(Vz - V).max(), "max|var-var_ref|", np.abs(var_oz - var_ref).max(), "min var", var_ref.min())
