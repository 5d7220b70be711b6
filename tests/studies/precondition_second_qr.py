import sys, numpy as np, pickle, scipy.linalg as sl
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
def stats(name, X):
    Y, hist = dm.jacobi_rows(X)
    n = X.shape[0]
    print(f'  {name:28s} sweeps {len(hist)} rot/pairs per sweep ' + ' '.join(f'{h/(n*(n-1)/2):.2f}' for h in hist), ' total', f'{sum(hist)/(n*(n-1)/2):.2f}')
    return Y
for (theta, chiR) in hv[:6]:
    perm = dm.interleave_perm(chiR)
    X = theta[:, perm]
    s = np.linalg.svd(theta, compute_uv=False)
    print('theta', theta.shape, 'sigma range', s[0], s[len(s)//2], s[-1])
    R = np.linalg.qr(X, mode='r')
    stats('A: QR(interleaved)', R)
    # B: second QR
    R1 = np.linalg.qr(R.conj().T, mode='r')
    stats('B: + second QR (L=R1^H)', R1.conj().T)
    # C: pivoted QR then rows
    Rp = sl.qr(theta, mode='r', pivoting=True)[0]
    stats('C: pivoted QR', Rp)
    R1p = np.linalg.qr(Rp.conj().T, mode='r')
    stats('D: pivoted + second QR', R1p.conj().T)
    R2 = np.linalg.qr(R1.conj().T, mode='r')
    stats('E: three QRs (rows of R2)', R2)  # R2 = Q2^H R1^H: left transform of L... 
