import sys, numpy as np, pickle
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
EPS = dm.EPS
def rot_pairs(X, nrm2, I, J, tol):
    ai, aj = nrm2[I], nrm2[J]
    g = np.sum(X[I]*X[J].conj(), axis=1)
    g2 = np.abs(g)**2
    act = g2 > tol*tol*ai*aj
    if not act.any(): return 0, 0
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
    return int(act.sum()), nbig
def jacobi_blocks(X, BR=16, p_order='asc', q_order='asc', internal='first', max_sweeps=48):
    X = np.array(X, dtype=complex); M, N = X.shape
    tol = 2*np.sqrt(N)*EPS
    nblk = M // BR
    hist = []
    for sw in range(max_sweeps):
        nrm2 = np.sum(np.abs(X)**2, axis=1)
        nrot = nbig = 0
        ps = range(nblk) if p_order == 'asc' else range(nblk-1, -1, -1)
        def do_internal(p):
            nonlocal nrot, nbig
            for r in range(BR-1):
                I, J = dm.rr_pairs(BR, r)
                a, b = rot_pairs(X, nrm2, I + p*BR, J + p*BR, tol); nrot += a; nbig += b
        for p in ps:
            if internal == 'first': do_internal(p)
            others = [q for q in range(nblk) if (q > p if p_order == 'asc' else q < p)]
            if q_order == 'desc': others = others[::-1]
            for q in others:
                w = np.arange(BR)
                for s in range(BR):
                    a, b = rot_pairs(X, nrm2, p*BR + w, q*BR + ((w+s) % BR), tol); nrot += a; nbig += b
            if internal == 'last': do_internal(p)
        hist.append(nrot)
        if nbig == 0: break
    return X, hist
res = {}
for (theta, chiR) in hv[:8]:
    perm = dm.interleave_perm(chiR)
    R = np.linalg.qr(theta[:, perm], mode='r')
    n = R.shape[0]; P = n*(n-1)/2
    for name, kw in [('device', {}), ('q desc', dict(q_order='desc')), ('p desc', dict(p_order='desc')), ('p desc q desc', dict(p_order='desc', q_order='desc')), ('internal last', dict(internal='last')), ('BR8', dict(BR=8)), ('BR32', dict(BR=32))]:
        Y, hist = jacobi_blocks(R, **kw)
        r = res.setdefault(name, [0, 0.0]); r[0] += len(hist); r[1] += sum(hist)/P
k = 8
for name, r in res.items(): print(f'{name:16s} mean sweeps {r[0]/k:.2f} mean rotations/pairs {r[1]/k:.2f}')
