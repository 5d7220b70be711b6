"""Harvest theta matrices from a device-model TEBD run for preconditioning studies."""
import sys, numpy as np, pickle
sys.path.insert(0, '.')
from oracle import tebd_ref, device_model as dm
L, chi = 16, 64
rng = np.random.RandomState(1000)
h = rng.uniform(-0.3, 0.3, L)
def run(psi, eps, nper, harvest=None):
    kick, gates = tebd_ref.make_gates(L, 1.0, h, 1.0, eps)
    for t in range(nper):
        for half in range(2):
            for start in (0, 1):
                for i in range(start, L - 1, 2):
                    if harvest is not None:
                        B0, B1 = psi.get_B(i, 'B'), psi.get_B(i + 1, 'B')
                        chiL, chiR = B0.shape[0], B1.shape[2]
                        C = np.tensordot(B0, B1, axes=(2, 0))
                        C = np.einsum('pqrs,arsb->apqb', np.asarray(gates[i]).reshape(2, 2, 2, 2), C).reshape(2 * chiL, 2 * chiR)
                        theta = C * np.repeat(psi._S[i], 2)[:, None]
                        if min(theta.shape) == 2 * chi:
                            harvest.append((theta, chiR))
                    psi.update_bond_tebd(i, gates[i], chi_max=chi, svd_min=1e-12, trunc_cut=1e-7)
            if half == 0:
                for i in range(L):
                    psi.apply_local_op(i, kick, unitary=True)
    return psi
psi = tebd_ref.product_state(L, 'neel', 1)
psi = run(psi, 0.3, 8)
print('chi after prep', psi.chi)
hv = []
psi = run(psi, 0.1, 2, hv)
print(len(hv), 'matrices', hv[0][0].shape)
pickle.dump(hv, open('tests/studies/_thetas.pkl', 'wb'))
