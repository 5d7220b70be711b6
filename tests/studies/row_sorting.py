"""de Rijk-style row ordering for the device's block sweep order: rows sorted by decreasing norm (a) once after the QR,
(b) before every sweep, (c) increasing.  TEBD matrices from harvest_thetas.py, thresholds as on the device.
    python tests/studies/row_sorting.py
"""
import pickle, sys
import numpy as np
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
EPS = dm.EPS
THR = dm.THRESHOLDS

def rot_pairs(X, nrm2, I, J, tol2, small2):
    ai, aj = nrm2[I], nrm2[J]
    g = np.sum(X[I] * X[J].conj(), axis=1)
    g2 = np.abs(g) ** 2
    nbig = int(np.sum(g2 > small2 * ai * aj))
    act = g2 > tol2 * ai * aj
    if not act.any():
        return 0, nbig
    I, J, g, ai, aj = I[act], J[act], g[act], ai[act], aj[act]
    ga = np.abs(g)
    dd = aj - ai
    t = np.copysign(2 * ga / (np.abs(dd) + np.sqrt(dd * dd + 4 * ga * ga)), dd)
    cs = 1 / np.sqrt(1 + t * t)
    se = cs * t * g / ga
    xi, xj = X[I], X[J]
    X[I] = cs[:, None] * xi - se[:, None] * xj
    X[J] = se.conj()[:, None] * xi + cs[:, None] * xj
    nrm2[I], nrm2[J] = ai - t * ga, aj + t * ga
    return int(act.sum()), nbig

def jacobi(X, BR=16, sort=None, small_rel2=1e-16, max_sweeps=48):
    X = np.array(X, dtype=complex)
    M, N = X.shape
    tol2f = (2 * np.sqrt(N) * EPS) ** 2
    nblk = M // BR
    rots = 0
    for sw in range(max_sweeps):
        nrm2 = np.sum(np.abs(X) ** 2, axis=1)
        if sort == 'each' or (sort in ('once', 'once_inc') and sw == 0):
            o = np.argsort(-nrm2 if sort != 'once_inc' else nrm2, kind='stable')
            X, nrm2 = X[o], nrm2[o]
        tol2 = max(tol2f, THR[sw]) if sw < 4 else tol2f
        small2 = tol2f if tol2 > tol2f else max(tol2f, small_rel2)
        nbig = 0
        for p in range(nblk):
            for r in range(BR - 1):
                I, J = dm.rr_pairs(BR, r)
                a, b = rot_pairs(X, nrm2, I + p * BR, J + p * BR, tol2, small2); rots += a; nbig += b
            w = np.arange(BR)
            for q in range(p + 1, nblk):
                for s in range(BR):
                    a, b = rot_pairs(X, nrm2, p * BR + w, q * BR + ((w + s) % BR), tol2, small2); rots += a; nbig += b
        if nbig == 0:
            break
    G = X @ X.conj().T
    d = np.sqrt(np.diag(G).real)
    A = np.abs(G) / np.outer(d, d); np.fill_diagonal(A, 0)
    return sw + 1, rots / (M * (M - 1) / 2), A.max()

res = {}
for (theta, chiR) in hv[:10]:
    perm = dm.interleave_perm(chiR)
    R = np.linalg.qr(theta[:, perm], mode='r')
    for name, kw in [('device', {}), ('sorted once', dict(sort='once')), ('sorted each sweep', dict(sort='each')),
                     ('increasing once', dict(sort='once_inc')), ('stop at 1e-13', dict(small_rel2=1e-13)),
                     ('stop at 1e-12', dict(small_rel2=1e-12))]:
        sw, rpp, off = jacobi(R, **kw)
        r = res.setdefault(name, [0, 0.0, 0.0]); r[0] += sw; r[1] += rpp; r[2] = max(r[2], off)
k = 10
for name, r in res.items():
    print(f'{name:20s} mean sweeps {r[0]/k:.2f}  rotations/pair {r[1]/k:.2f}  worst residual |g|/sqrt(aa) {r[2]:.1e}')
