"""
O1 -- exact state-vector simulator of the reference's Floquet sequence.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``src/models/kicked_ising.py:73-160`` of the reference: the same 4x4 bond gates
(each bulk site receives its field in both adjacent bond gates), even bonds,
odd bonds, kick on every site, even bonds, odd bonds.  No truncation, so it is
ground truth whenever the MPS bond dimension is not limited
(chi_cap >= 2**(L//2)).  L <= 20 in practice.

Observables follow ``src/core/observables.py:11-71`` (raw sigma_z = diag(+1,-1)
on the internal basis index) and TeNPy's entropy definition.
"""

import numpy as np

from .tebd_ref import make_gates


def basis_state(n_sites, state='neel', up_index=1):
    """Amplitudes psi[p0,...,p_{L-1}] of the product states of tensor_utils.py:44-55."""
    if state == 'all_up':
        idx = [up_index] * n_sites
    elif state == 'all_down':
        idx = [1 - up_index] * n_sites
    elif state == 'neel':
        idx = [up_index if i % 2 == 0 else 1 - up_index for i in range(n_sites)]
    else:
        raise ValueError(f"Unknown state type: {state}")
    psi = np.zeros((2,) * n_sites, dtype=complex)
    psi[tuple(idx)] = 1.0
    return psi


def apply_two_site(psi, gate, i):
    """psi[..., p_i, p_{i+1}, ...] <- gate[(p_i p_{i+1}), (q_i q_{i+1})] psi[..., q_i, q_{i+1}, ...]."""
    L = psi.ndim
    g = np.asarray(gate).reshape(2, 2, 2, 2)
    psi = np.tensordot(g, psi, axes=((2, 3), (i, i + 1)))      # (p_i, p_{i+1}, rest...)
    return np.moveaxis(psi, (0, 1), (i, i + 1))


def apply_one_site(psi, op, i):
    psi = np.tensordot(op, psi, axes=((1,), (i,)))
    return np.moveaxis(psi, 0, i)


def floquet_step(psi, kick, gates):
    L = psi.ndim
    for _half in range(2):
        for i in range(0, L - 1, 2):
            psi = apply_two_site(psi, gates[i], i)
        for i in range(1, L - 1, 2):
            psi = apply_two_site(psi, gates[i], i)
        if _half == 0:
            for i in range(L):
                psi = apply_one_site(psi, kick, i)
    return psi


def site_z(psi):
    L = psi.ndim
    p = np.abs(psi) ** 2
    out = np.empty(L)
    for i in range(L):
        pi = p.sum(axis=tuple(j for j in range(L) if j != i))
        out[i] = pi[0] - pi[1]
    return out


def schmidt_values(psi, bond):
    """Singular values across the cut between sites bond-1 and bond."""
    L = psi.ndim
    return np.linalg.svd(psi.reshape(2 ** bond, 2 ** (L - bond)), compute_uv=False)


def bond_entropies(psi):
    L = psi.ndim
    out = []
    for b in range(1, L):
        p = schmidt_values(psi, b) ** 2
        p = p[p > 1e-30]
        out.append(-np.sum(p * np.log(p)))
    return np.array(out)


def run(n_sites, J, h_fields, tau, n_periods, epsilon=0.0, state='neel', up_index=1,
        entropies=True):
    """Returns dict Z[t,i], S_ent[t,b], LE[t], times[t] for t = 0..n_periods."""
    kick, gates = make_gates(n_sites, J, h_fields, tau, epsilon)
    psi0 = basis_state(n_sites, state, up_index)
    psi = psi0.copy()
    Z, Sent, LE, times = [], [], [], []
    for t in range(n_periods + 1):
        if t > 0:
            psi = floquet_step(psi, kick, gates)
        Z.append(site_z(psi))
        if entropies:
            Sent.append(bond_entropies(psi))
        LE.append(abs(np.vdot(psi0, psi)) ** 2)
        times.append(t * 2 * tau)
    return dict(Z=np.array(Z), S_ent=np.array(Sent), LE=np.array(LE), times=np.array(times), psi=psi)
