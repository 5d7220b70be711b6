import sys, numpy as np, pickle, scipy.linalg as sl
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
EPS = dm.EPS
def jacobi_rows_abs(X, tol_abs, small_abs, max_sweeps=48):
    X = np.array(X, dtype=complex); M, N = X.shape
    tol = 2*np.sqrt(N)*EPS
    hist = []
    tot = np.sum(np.abs(X)**2)
    amax = np.max(np.sum(np.abs(X)**2, axis=1))
    for _ in range(max_sweeps):
        nrm2 = np.sum(np.abs(X)**2, axis=1)
        nrot = nbig = 0
        for r in range(M-1):
            I, J = dm.rr_pairs(M, r)
            ai, aj = nrm2[I], nrm2[J]
            g = np.sum(X[I]*X[J].conj(), axis=1)
            g2 = np.abs(g)**2
            act = (g2 > tol*tol*ai*aj) & (g2 > tol_abs**2 * amax*amax)
            if not act.any(): continue
            I, J, g, ai, aj = I[act], J[act], g[act], ai[act], aj[act]
            ga = np.abs(g)
            nbig += int(np.sum((ga*ga > 1e-16*ai*aj) & (ga*ga > small_abs**2*amax*amax)))
            dd = aj-ai
            t = np.copysign(2*ga/(np.abs(dd)+np.sqrt(dd*dd+4*ga*ga)), dd)
            cs = 1/np.sqrt(1+t*t); se = cs*t*g/ga
            xi, xj = X[I], X[J]
            X[I] = cs[:,None]*xi - se[:,None]*xj
            X[J] = se.conj()[:,None]*xi + cs[:,None]*xj
            nrm2[I], nrm2[J] = ai - t*ga, aj + t*ga
            nrot += int(act.sum())
        hist.append(nrot)
        if nbig == 0: break
    return X, hist
for (theta, chiR) in hv[:4]:
    perm = dm.interleave_perm(chiR)
    R = np.linalg.qr(theta[:, perm], mode='r')
    n = R.shape[0]; k = n//2
    U, s, Vh = np.linalg.svd(R)
    for tol_abs, small_abs in [(0, 0), (1e-16, 1e-12), (1e-15, 1e-11), (1e-14, 1e-10), (1e-13, 1e-9)]:
        Y, hist = jacobi_rows_abs(R, tol_abs, small_abs)
        w = np.linalg.norm(Y, axis=1); o = np.argsort(-w); Yk = Y[o[:k]]; Vk = Yk/w[o[:k], None]
        # subspace error vs LAPACK, orthonormality, gram off-diagonal absolute
        P = Vk.conj().T @ Vk; Pl = Vh[:k].conj().T @ Vh[:k]
        G = Yk @ Yk.conj().T; off = G - np.diag(np.diag(G))
        print(f'tol_abs {tol_abs:7.0e} sweeps {len(hist)} total {sum(hist)/(n*(n-1)/2):.2f} | projector diff {np.linalg.norm(P-Pl):.2e} state diff {np.linalg.norm(R@(P-Pl)):.2e} orth {np.linalg.norm(Vk@Vk.conj().T-np.eye(k)):.2e} max|gamma| {np.abs(off).max():.2e} sum|gamma| {np.abs(off).sum():.2e} dsig {np.abs(w[o]-s).max():.2e}')
    print()
