import sys, numpy as np, pickle
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
EPS = dm.EPS
def rot_pairs(X, nrm2, I, J, tol2, thr2):
    ai, aj = nrm2[I], nrm2[J]
    g = np.sum(X[I]*X[J].conj(), axis=1)
    g2 = np.abs(g)**2
    above_tol = g2 > tol2*ai*aj
    act = g2 > max(tol2, thr2)*ai*aj
    pending = int(np.sum(above_tol & ~act))      # skipped although above the final tolerance
    if not act.any(): return 0, 0, pending
    I, J, g, ai, aj = I[act], J[act], g[act], ai[act], aj[act]
    ga = np.abs(g)
    nbig = int(np.sum(ga*ga > 1e-16*ai*aj))
    dd = aj-ai
    t = np.copysign(2*ga/(np.abs(dd)+np.sqrt(dd*dd+4*ga*ga)), dd)
    cs = 1/np.sqrt(1+t*t); se = cs*t*g/ga
    xi, xj = X[I], X[J]
    X[I] = cs[:,None]*xi - se[:,None]*xj
    X[J] = se.conj()[:,None]*xi + cs[:,None]*xj
    nrm2[I], nrm2[J] = ai - t*ga, aj + t*ga
    return int(act.sum()), nbig, pending
def jacobi_blocks(X, sched, BR=16, max_sweeps=48):
    X = np.array(X, dtype=complex); M, N = X.shape
    tol = 2*np.sqrt(N)*EPS; tol2 = tol*tol
    nblk = M // BR
    hist = []
    for sw in range(max_sweeps):
        thr2 = sched[sw] if sw < len(sched) else 0.0
        nrm2 = np.sum(np.abs(X)**2, axis=1)
        nrot = nbig = npend = 0
        for p in range(nblk):
            for r in range(BR-1):
                I, J = dm.rr_pairs(BR, r)
                a, b, c = rot_pairs(X, nrm2, I + p*BR, J + p*BR, tol2, thr2); nrot += a; nbig += b; npend += c
            for q in range(p+1, nblk):
                w = np.arange(BR)
                for s in range(BR):
                    a, b, c = rot_pairs(X, nrm2, p*BR + w, q*BR + ((w+s) % BR), tol2, thr2); nrot += a; nbig += b; npend += c
        hist.append(nrot)
        if nbig == 0 and npend == 0: break
    return X, hist
scheds = {'none': [], 'C': [1e-2,1e-3,1e-4,1e-6], 'F 1e-2,3e-3,1e-3,1e-4,1e-6': [1e-2,3e-3,1e-3,1e-4,1e-6], 'G 1e-2,1e-2,1e-3,1e-5': [1e-2,1e-2,1e-3,1e-5], 'H 1e-2,1e-3,1e-4,1e-5,1e-7': [1e-2,1e-3,1e-4,1e-5,1e-7], 'I 1e-2,1e-3,1e-5': [1e-2,1e-3,1e-5], 'J 1e-2,1e-3,1e-4,1e-6,1e-9': [1e-2,1e-3,1e-4,1e-6,1e-9]}
res = {}
for (theta, chiR) in hv[:12]:
    perm = dm.interleave_perm(chiR)
    R = np.linalg.qr(theta[:, perm], mode='r')
    n = R.shape[0]; P = n*(n-1)/2
    s_ref = np.linalg.svd(R, compute_uv=False)
    for name, sc in scheds.items():
        Y, hist = jacobi_blocks(R, sc)
        w = np.sort(np.linalg.norm(Y, axis=1))[::-1]
        r = res.setdefault(name, [0, 0.0, 0.0]); r[0] += len(hist); r[1] += sum(hist)/P; r[2] = max(r[2], np.max(np.abs(w-s_ref)[:n//2]))
k = 12
for name, r in res.items(): print(f'{name:24s} mean sweeps {r[0]/k:.2f} rotations/pairs {r[1]/k:.2f} cost(0.4 visit + 0.6 rot) {0.4*r[0]/k + 0.6*r[1]/k:.2f} max dsig {r[2]:.1e}')
