"""Sweep counts of the one-sided Jacobi on 512 x 512 two-site tensors of BASELINE config 4's regime
(tests/studies/harvest_wide_thetas.py): device order with and without the threshold schedule, columns sorted by norm
before the QR, column-pivoted QR."""
import sys, numpy as np, pickle, scipy.linalg as sl, time
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas_wide.pkl', 'rb'))
print(len(hv), 'matrices')
def stats(name, X, thresholds=dm.THRESHOLDS):
    t0=time.time()
    Y, hist = dm.jacobi_rows(X, thresholds=thresholds)
    n = X.shape[0]
    print(f'  {name:34s} sweeps {len(hist):2d}  rot/pairs per sweep ' + ' '.join(f'{h/(n*(n-1)/2):.2f}' for h in hist), ' total', f'{sum(hist)/(n*(n-1)/2):.2f}', f'({time.time()-t0:.0f}s)', flush=True)
    return Y
for (theta, chiR) in hv[:2]:
    perm = dm.interleave_perm(chiR)
    X = theta[:, perm]
    s = np.linalg.svd(theta, compute_uv=False)
    print('theta', theta.shape, 'sigma', s[0], s[len(s)//4], s[len(s)//2], s[-1], 'n(sigma<1e-10 s0):', int((s < 1e-10*s[0]).sum()))
    R = np.linalg.qr(X, mode='r')
    stats('A: QR(interleaved)', R)
    stats('A0: no thresholds', R, thresholds=())
    # column sort by norm before the QR
    cn = np.linalg.norm(X, axis=0)
    Rs = np.linalg.qr(X[:, np.argsort(-cn)], mode='r')
    stats('S: columns sorted by norm, QR', Rs)
    Rp = sl.qr(theta, mode='r', pivoting=True)[0]
    stats('C: pivoted QR', Rp)
