import sys, numpy as np, pickle
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
EPS = dm.EPS
def jacobi_freeze(X, chi_max, f, freeze_after, max_sweeps=48):
    X = np.array(X, dtype=complex); M, N = X.shape
    tol = 2*np.sqrt(N)*EPS
    hist = []; visited = []
    frozen = np.zeros(M, bool)
    for sw in range(max_sweeps):
        nrm2 = np.sum(np.abs(X)**2, axis=1)
        if sw == freeze_after:
            a_cut = np.sort(nrm2)[::-1][chi_max-1]
            frozen = nrm2 < f*a_cut
        nrot = nbig = nvis = 0
        for r in range(M-1):
            I, J = dm.rr_pairs(M, r)
            keep = ~(frozen[I] & frozen[J])
            I, J = I[keep], J[keep]
            nvis += len(I)
            ai, aj = nrm2[I], nrm2[J]
            g = np.sum(X[I]*X[J].conj(), axis=1)
            g2 = np.abs(g)**2
            act = g2 > tol*tol*ai*aj
            if not act.any(): continue
            I, J, g, ai, aj = I[act], J[act], g[act], ai[act], aj[act]
            ga = np.abs(g)
            nbig += int(np.sum(ga*ga > 1e-16*ai*aj))
            dd = aj-ai
            t = np.copysign(2*ga/(np.abs(dd)+np.sqrt(dd*dd+4*ga*ga)), dd)
            cs = 1/np.sqrt(1+t*t); se = cs*t*g/ga
            xi, xj = X[I], X[J]
            X[I] = cs[:,None]*xi - se[:,None]*xj
            X[J] = se.conj()[:,None]*xi + cs[:,None]*xj
            nrm2[I], nrm2[J] = ai - t*ga, aj + t*ga
            nrot += int(act.sum())
        hist.append(nrot); visited.append(nvis)
        if nbig == 0: break
    return X, hist, visited, frozen
tot = {}
for (theta, chiR) in hv:
    perm = dm.interleave_perm(chiR)
    R = np.linalg.qr(theta[:, perm], mode='r')
    n = R.shape[0]; k = n//2; P = n*(n-1)/2
    U, s, Vh = np.linalg.svd(R)
    for fa, f in ((99, 0), (4, 0.25), (4, 0.05), (3, 0.05), (4, 0.01), (5, 0.25)):
        Y, hist, vis, frozen = jacobi_freeze(R, k, f, fa)
        w = np.linalg.norm(Y, axis=1); o = np.argsort(-w); Yk = Y[o[:k]]; Vk = Yk/w[o[:k], None]
        Pk = Vk.conj().T @ Vk; Pl = Vh[:k].conj().T @ Vh[:k]
        cost = sum(0.3*v + 0.7*h for v, h in zip(vis, hist))/P
        key = (fa, f)
        t = tot.setdefault(key, [0, 0, 0, 0, 0])
        t[0] += cost; t[1] += len(hist); t[2] = max(t[2], np.linalg.norm(Pk-Pl)); t[3] = max(t[3], np.abs(w[o[:k]]-s[:k]).max()); t[4] += frozen.sum()
n = len(hv)
for key, t in tot.items():
    print(f'freeze_after {key[0]:2d} f {key[1]:5.2f}: mean cost {t[0]/n:.2f} mean sweeps {t[1]/n:.2f} max proj diff {t[2]:.1e} max dsig kept {t[3]:.1e} mean frozen {t[4]/n:.0f}')
