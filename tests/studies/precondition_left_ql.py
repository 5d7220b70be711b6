import sys, numpy as np, pickle, scipy.linalg as sl
sys.path.insert(0, '.')
from oracle import device_model as dm
hv = pickle.load(open('tests/studies/_thetas.pkl', 'rb'))
def ql(A):
    # A = Q L  via QR of the flipped matrix
    n = A.shape[0]
    J = np.eye(n)[::-1]
    Q, R = np.linalg.qr(A[::-1, ::-1])   # A J-flipped
    # A = J (A_flip) J ; A_flip = Q R -> A = (J Q J)(J R J), J R J is lower triangular
    return Q[::-1, ::-1], R[::-1, ::-1]
def stats(name, X, Vref=None):
    Y, hist = dm.jacobi_rows(X)
    n = X.shape[0]
    print(f'  {name:34s} sweeps {len(hist)} rot/pairs ' + ' '.join(f'{h/(n*(n-1)/2):.2f}' for h in hist), ' total', f'{sum(hist)/(n*(n-1)/2):.2f}')
    return Y
for (theta, chiR) in hv[:4]:
    perm = dm.interleave_perm(chiR)
    X = theta[:, perm]
    R = np.linalg.qr(X, mode='r')
    stats('A: rows of R (current)', R)
    Q3, L3 = ql(R)
    assert np.allclose(Q3 @ L3, R)
    stats('F: rows of L3 (QL of R)', L3)
    R4 = np.linalg.qr(L3, mode='r')
    stats('G: rows of R4 (QR of L3)', R4)
    Q5, L5 = ql(R4)
    stats('H: rows of L5 (QL of R4)', L5)
    # sorted rows of R by norm descending then QR? (left permutation is free)
    print()
