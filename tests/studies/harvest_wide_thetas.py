"""Harvest 512 x 512 two-site tensors in BASELINE config 4's regime (eps = 0.1 from the Neel state, W = 0.3,
chi_max = 256) for the sweep-count study of the wide matrices: L = 24 instead of 64 (the centre bonds saturate the same
way), evolved on the CPU oracle until the middle bond has reached chi_max, then two more periods harvested.
-> tests/studies/_thetas_wide.pkl (git-ignored).  Takes several minutes."""
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from oracle import tebd_ref  # noqa: E402

L, chi, eps = 24, 256, 0.1
rng = np.random.RandomState(1000)
h = rng.uniform(-0.3, 0.3, L)
kick, gates = tebd_ref.make_gates(L, 1.0, h, 1.0, eps)


def period(psi, harvest=None):
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


psi = tebd_ref.product_state(L, 'neel', 1)
t0 = time.time()
t = 0
while psi.chi[L // 2 - 1] < chi and t < 120:
    period(psi)
    t += 1
    if t % 5 == 0:
        print(t, 'chi_mid', psi.chi[L // 2 - 1], round(time.time() - t0, 1), 's', flush=True)
hv = []
period(psi, hv)
period(psi, hv)
print(len(hv), 'matrices after', t, 'periods', flush=True)
pickle.dump(hv[:: max(1, len(hv) // 8)][:8], open('tests/studies/_thetas_wide.pkl', 'wb'))
