"""Block one-sided Jacobi with a Gram eigen-solve per block pair (the structure VERDICT r01 item 3 names):
for every pair of row blocks (P, Q) of size b: G = X_PQ X_PQ^H (2b x 2b, a GEMM), G = W Lambda W^H, X_PQ <- W^H X_PQ
(a GEMM).  With an exact inner solve the sweep count is the best any inner solver can reach; `inner_sweeps` = k models
k cyclic two-sided Jacobi sweeps on G instead.  Compared with the device's scalar sweep order on the same TEBD matrices
(128 x 128 from harvest_thetas.py), same stopping rule (a sweep that found every pair below 1e-8 relative ends it).

    python tests/studies/block_gram_jacobi.py
"""
import pickle
import sys

import numpy as np

sys.path.insert(0, '.')
from oracle import device_model as dm

hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
EPS = dm.EPS


def offdiag_rel(G):
    d = np.sqrt(np.abs(np.diag(G).real))
    A = np.abs(G) / np.maximum(np.outer(d, d), 1e-300)
    np.fill_diagonal(A, 0.0)
    return A.max()


def inner_jacobi(G, sweeps):
    """k cyclic two-sided Jacobi sweeps on the Hermitian G; returns the accumulated unitary W (G ~ W D W^H)."""
    n = G.shape[0]
    G = G.copy()
    W = np.eye(n, dtype=complex)
    for _ in range(sweeps):
        for i in range(n - 1):
            for j in range(i + 1, n):
                g = G[i, j]
                if abs(g) ** 2 <= (EPS ** 2) * abs(G[i, i].real * G[j, j].real):
                    continue
                dd = G[j, j].real - G[i, i].real
                t = np.copysign(2 * abs(g) / (abs(dd) + np.sqrt(dd * dd + 4 * abs(g) ** 2)), dd)
                c = 1 / np.sqrt(1 + t * t)
                se = c * t * g / abs(g)
                J = np.eye(n, dtype=complex)
                J[i, i] = J[j, j] = c
                J[i, j] = se          # rows: x_i' = c x_i - se x_j  <=>  X' = J^H X with this J
                J[j, i] = -np.conj(se)
                G = J.conj().T @ G @ J
                W = W @ J
    return W


def block_jacobi(X, b, inner_sweeps=None, max_sweeps=30):
    X = np.array(X, dtype=complex)
    M, N = X.shape
    nblk = M // b
    visits = 0
    for sw in range(max_sweeps):
        worst = 0.0
        for p in range(nblk):
            for q in range(p + 1, nblk):
                idx = np.r_[p * b:(p + 1) * b, q * b:(q + 1) * b]
                Y = X[idx]
                G = Y @ Y.conj().T
                worst = max(worst, offdiag_rel(G))
                if inner_sweeps is None:
                    lam, W = np.linalg.eigh(G)
                    W = W[:, ::-1]
                else:
                    W = inner_jacobi(G, inner_sweeps)
                X[idx] = W.conj().T @ Y
                visits += 1
        if worst < 1e-8:
            return X, sw + 1
    return X, max_sweeps


def accuracy(X, theta_p):
    s = np.sort(np.linalg.norm(X, axis=1))[::-1]
    s_ref = np.linalg.svd(theta_p, compute_uv=False)
    k = min(len(s), len(s_ref))
    rel = np.max(np.abs(s[:k] - s_ref[:k]) / np.maximum(s_ref[:k], 1e-300) * (s_ref[:k] > 1e-13 * s_ref[0]))
    G = X @ X.conj().T
    return rel, offdiag_rel(G)


rows = []
for (theta, chiR) in hv[:6]:
    perm = dm.interleave_perm(chiR)
    R = np.linalg.qr(theta[:, perm], mode='r')
    Y, hist = dm.jacobi_rows(R, thresholds=dm.THRESHOLDS) if False else (None, None)
    res = {}
    for name, kw in [('b=8 exact', dict(b=8)), ('b=16 exact', dict(b=16)), ('b=32 exact', dict(b=32)),
                     ('b=16 inner 1 sweep', dict(b=16, inner_sweeps=1)), ('b=16 inner 2 sweeps', dict(b=16, inner_sweeps=2))]:
        Xb, sw = block_jacobi(R, **kw)
        rel, off = accuracy(Xb, theta[:, perm])
        res[name] = (sw, rel, off)
    rows.append(res)
    print({k: (v[0], f'{v[1]:.1e}', f'{v[2]:.1e}') for k, v in res.items()}, flush=True)
print()
for name in rows[0]:
    print(f'{name:22s} mean sweeps {np.mean([r[name][0] for r in rows]):.2f}  worst sigma rel err '
          f'{max(r[name][1] for r in rows):.1e}  worst residual off-diagonal {max(r[name][2] for r in rows):.1e}')
