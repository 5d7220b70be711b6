"""Re-canonicalisation after imaginary-time TEBD by sweeps of identity gates (NumPy model, CPU only).

TFIM, L = 12, chi_max = 16: after 100 second-order steps of exp(-dt H) the chain is far from canonical form
(max |sum_p B_p B_p^+ - 1| = 0.7 at dt = 0.1, 0.2 at dt = 0.02) and the energy read off the canonical-form two-site
expectation values differs from <psi|H|psi> / <psi|psi> of the same tensors by 4e-3 / 8e-4.  Even + odd layers of
identity gates without truncation leave the state (the product of the B tensors) untouched and restore the canonical
form: error 0.12, 0.04, 5e-3, 3e-4, 7e-5 after sweeps 1..5 and exactly (1e-15) after L/2 = 6; the canonical-form
energy then equals the exact expectation value to 1e-13.  This is what MPS.canonical_form does on the device.
Run from the repository root: python tests/studies/recanonicalise_identity_layers.py
"""
import sys, numpy as np, scipy.linalg as sl
sys.path.insert(0, '.')
from oracle import tebd_ref
L, g = 12, 0.7
X, Z, I2 = tebd_ref.SIGMA_X, tebd_ref.SIGMA_Z, np.eye(2)
Hb = []
for i in range(L - 1):
    wl = 1.0 if i == 0 else 0.5
    wr = 1.0 if i == L - 2 else 0.5
    Hb.append(-np.kron(Z, Z) - g * (wl * np.kron(X, I2) + wr * np.kron(I2, X)))
H = np.zeros((2 ** L, 2 ** L), dtype=complex)
for i, hb in enumerate(Hb):
    H += np.kron(np.kron(np.eye(2 ** i), hb), np.eye(2 ** (L - i - 2)))
e0 = np.linalg.eigvalsh(H)[0]
psi = tebd_ref.product_state(L, 'all_up', 1)
rot = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
for i in range(L):
    psi.apply_local_op(i, rot, unitary=True)
def layer(psi, gates, par, chi=16):
    for i in range(par, L - 1, 2):
        psi.update_bond_tebd(i, gates[i], chi_max=chi, svd_min=1e-14)
def canon_err(psi):
    eL = eR = 0.0
    for i in range(L):
        B = psi.get_B(i, 'B')
        R = np.einsum('apb,cpb->ac', B, B.conj())
        eR = max(eR, np.abs(R - np.eye(R.shape[0])).max())
        A = psi.get_B(i, 'A')
        Lm = np.einsum('apb,apc->bc', A.conj(), A)
        eL = max(eL, np.abs(Lm - np.eye(Lm.shape[0])).max())
    return eL, eR
def e_local(psi):
    tot = 0.0
    for b, h in enumerate(Hb):
        th = psi.get_theta(b, 2)   # (vL,p0,p1,vR)
        tot += np.einsum('apqb,pqrs,arsb->', th.conj(), h.reshape(2,2,2,2), th).real
    return tot
def e_exact(psi):
    v = psi.to_statevector().reshape(-1)
    return (v.conj() @ H @ v).real / (v.conj() @ v).real
for dt in (0.1, 0.02):
    ge = [sl.expm(-dt * (0.5 if b % 2 == 0 else 1.0) * h) for b, h in enumerate(Hb)]
    for s in range(100):
        layer(psi, ge, 0); layer(psi, ge, 1); layer(psi, ge, 0)
    print('dt', dt, 'canon err', canon_err(psi), 'E_local - e0', e_local(psi) - e0, 'E_exact - e0', e_exact(psi) - e0)
ident = [np.eye(4)] * (L - 1)
for k in range(8):
    layer(psi, ident, 0, chi=None); layer(psi, ident, 1, chi=None)
    print('identity sweep', k + 1, 'canon err', canon_err(psi), 'E_local - e0', e_local(psi) - e0, 'E_exact - e0', e_exact(psi) - e0, 'chi', max(psi.chi))
