"""
NumPy model of the DEVICE SVD algorithm, stage by stage -- TEST INFRASTRUCTURE.

The CUDA path does not call LAPACK.  Per two-site update it runs (tc_engine.cu / tc_jacobi*.cuh)

  K1   C      = gate . (kick B_i)(kick B_{i+1})          no left Schmidt values
       X      = diag(S_i (x) 1_2) C P                     theta with columns in interleaved order 2 b + p1
  K2a  X      = Q R,  Q discarded                        Householder QR from the left (zgeqr2 convention)
  K2b  J^H R  = Sigma V^H P                              one-sided Jacobi on the ROWS of R (round-robin order)
  K2c  sort rows by norm, truncate, renormalise;  B_{i+1} = V_k^H (columns un-permuted)
  K3   B_i    = C V_k / |Sigma_k|                        inverse-free update

This file restates K2a-K2c in plain NumPy so that (i) the choice of algorithm can be studied on the CPU
(sweep counts with and without the QR preconditioner, accuracy against LAPACK zgesdd on matrices
taken from TEBD runs) and (ii) the GPU tests have a stage-wise model for the workspace buffers
exposed by tc_dbg_get.

History: a Gram-matrix eigensolver route was studied first and dropped: it loses a factor 1/sigma_k of
accuracy in the vectors of small kept singular values and drifts from the LAPACK oracle by ~1e-7 over
60 periods at chi_max = 32, i.e. misses the 1e-8 parity bar.
"""

import numpy as np

from . import tebd_ref

EPS = np.finfo(float).eps
DEAD_REL2 = 1e-30          # rows with |x|^2 < DEAD_REL2 |theta|_F^2 are numerically zero (tc_jacobi.cuh)
MAX_SWEEPS = 48
SMALL_REL2 = 1e-16       # a sweep whose rotations were all below 1e-8 relative ends the iteration
THRESHOLDS = (3e-3, 3e-4, 3e-5, 3e-6)   # tc_engine.cu thr_sched: sweeps 0..3 rotate only pairs with |g|^2/(a_i a_j) above


def interleave_perm(chi_r):
    """Column order of X: position 2 b + p1 holds column (p1, b) = p1 * chi_r + b of theta."""
    return np.array([p * chi_r + b for b in range(chi_r) for p in range(2)])


def larfg(alpha, x):
    """LAPACK zlarfg: (beta, tau, v_tail) with H = I - tau [1;v][1;v]^H, H^H [alpha; x] = [beta; 0]."""
    xnorm = np.linalg.norm(x)
    if xnorm == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0 + 0.0j, np.zeros_like(x)
    beta = -np.copysign(np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm ** 2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    return beta, tau, x / (alpha - beta)


def householder_r(X):
    """K2a: the triangular factor of X (M x N), K = min(M, N) rows, reflectors discarded."""
    A = np.array(X, dtype=complex)
    M, N = A.shape
    for k in range(min(M - 1, N)):
        beta, tau, vt = larfg(A[k, k], A[k + 1:, k])
        if tau == 0:
            continue
        v = np.concatenate(([1.0 + 0j], vt))
        A[k, k] = beta
        A[k + 1:, k] = 0.0
        blk = A[k:, k + 1:]
        blk -= np.outer(v, np.conj(tau) * (v.conj() @ blk))
    return A[:min(M, N)]


def rr_pairs(M, r):
    """Round r of the circle-method tournament among M (even) rows: M/2 disjoint pairs (i < j)."""
    m1 = M - 1
    I, J = [], []
    for k in range(M // 2):
        if k == 0:
            i, j = m1, r
        else:
            i, j = (r + k) % m1, (r - k) % m1
        if i > j:
            i, j = j, i
        I.append(i)
        J.append(j)
    return np.array(I), np.array(J)


def jacobi_rows(X, tol=None, max_sweeps=MAX_SWEEPS, thresholds=()):
    """K2b: one-sided Jacobi on rows.  Returns (X with mutually orthogonal rows, rotations per sweep).
    A pair rotates when |g|^2 > tol^2 a_i a_j, g = x_i . conj(x_j); rotation
    x_i' = c x_i - (s e) x_j, x_j' = conj(s e) x_i + c x_j with e = g/|g| (same formulas as the kernels).
    ``thresholds`` (the 16-warp kernel uses THRESHOLDS): sweep k < len(thresholds) rotates only the pairs with
    |g|^2 > thresholds[k] a_i a_j; such a sweep cannot end the iteration while any pair is above the tolerance."""
    X = np.array(X, dtype=complex)
    M, N = X.shape
    if M % 2:
        raise ValueError('even number of rows expected (M = 2 chi)')
    if tol is None:
        tol = 2 * np.sqrt(N) * EPS
    dead = DEAD_REL2 * np.sum(np.abs(X) ** 2)
    hist = []
    for _sweep in range(max_sweeps):
        nrm2 = np.sum(np.abs(X) ** 2, axis=1)
        nrot = nbig = 0
        thr2 = max(tol * tol, thresholds[_sweep]) if _sweep < len(thresholds) else tol * tol
        small2 = tol * tol if thr2 > tol * tol else SMALL_REL2
        for r in range(M - 1):
            I, J = rr_pairs(M, r)
            ai, aj = nrm2[I], nrm2[J]
            g = np.sum(X[I] * X[J].conj(), axis=1)
            alive = (ai > dead) & (aj > dead)
            nbig += int(np.sum(alive & (np.abs(g) ** 2 > small2 * ai * aj)))
            act = alive & (np.abs(g) ** 2 > thr2 * ai * aj)
            if not act.any():
                continue
            I, J, g, ai, aj = I[act], J[act], g[act], ai[act], aj[act]
            ga = np.abs(g)
            dd = aj - ai
            t = np.copysign(2 * ga / (np.abs(dd) + np.sqrt(dd * dd + 4 * ga * ga)), dd)
            cs = 1 / np.sqrt(1 + t * t)
            se = cs * t * g / ga
            xi, xj = X[I], X[J]
            X[I] = cs[:, None] * xi - se[:, None] * xj
            X[J] = se.conj()[:, None] * xi + cs[:, None] * xj
            nrm2[I], nrm2[J] = ai - t * ga, aj + t * ga
            nrot += int(act.sum())
        hist.append(nrot)
        if nbig == 0:       # only rotations below 1e-8 relative (or none): converged (tc_jacobi.cuh)
            break
    return X, hist


def truncate_device(sig_sorted, mode, cutoff=1e-13, chi_max=None, svd_min=None, trunc_cut=None, chi_cap=None):
    """K2c truncation rule on singular values sorted descending (finalize_kernel).
    Returns (k, S_new, renorm, discarded weight / total weight)."""
    sig = np.asarray(sig_sorted, dtype=float)
    n = len(sig)
    tot2 = float(np.sum(sig[::-1] ** 2))
    tot = np.sqrt(tot2)
    if mode == 'reference':
        k = int(np.sum(sig > cutoff))
    else:
        # TeNPy truncate() on the normalised spectrum written for a descending array: every rule is a
        # maximal keep-count; a rule that would keep nothing is ignored.
        k = n if not chi_max else min(n, chi_max)
        k_svd = int(np.sum(sig > (svd_min or 0.0) * tot))
        if 1 <= k_svd < k:
            k = k_svd
        if trunc_cut:
            c = np.cumsum(sig[::-1] ** 2)
            good = c > trunc_cut * trunc_cut * tot2
            if good.any():
                k = min(k, n - int(np.argmax(good)))
    k = max(k, 1)
    if chi_cap is not None:
        k = min(k, chi_cap)
    renorm = np.linalg.norm(sig[:k][::-1])
    return k, sig[:k] / renorm, renorm, float(np.sum(sig[k:] ** 2) / tot2) if tot2 > 0 else 0.0


def svd_right(theta, chi_r=None, precondition=True):
    """K2a + K2b on theta (M x N): returns (sigma desc, Vh rows = right singular vectors in theta's own
    column order, info)."""
    theta = np.asarray(theta, dtype=complex)
    M, N = theta.shape
    chi_r = chi_r or N // 2
    perm = interleave_perm(chi_r) if precondition else np.arange(N)
    X = theta[:, perm]
    R = householder_r(X) if precondition else X
    Y, hist = jacobi_rows(R)
    w = np.linalg.norm(Y, axis=1)
    order = np.argsort(-w, kind='stable')
    w = w[order]
    Vh = np.zeros((len(order), N), dtype=complex)
    nz = w > 0
    Vh[np.ix_(nz, perm)] = Y[order][nz] / w[nz, None]
    return w, Vh, dict(sweeps=len(hist), rotations=hist)


def update_bond_device(psi, i, gate, mode='reference', trunc=None, chi_cap=None, precondition=True):
    """One two-site update with the device algorithm on an all-'B'-form oracle MPS."""
    trunc = trunc or {}
    B0, B1 = psi.get_B(i, 'B'), psi.get_B(i + 1, 'B')
    chiL, chiR = B0.shape[0], B1.shape[2]
    C = np.tensordot(B0, B1, axes=(2, 0))
    C = np.einsum('pqrs,arsb->apqb', np.asarray(gate).reshape(2, 2, 2, 2), C).reshape(2 * chiL, 2 * chiR)
    theta = C * np.repeat(psi._S[i], 2)[:, None]
    sig, Vh, info = svd_right(theta, chiR, precondition)
    k, S_new, renorm, err = truncate_device(sig, mode, chi_cap=chi_cap, **trunc)
    Vk = Vh[:k]
    psi._B[i + 1] = Vk.reshape(k, 2, chiR)
    psi._B[i] = (C @ Vk.conj().T / renorm).reshape(chiL, 2, k)
    psi.form[i] = psi.form[i + 1] = (0.0, 1.0)
    psi._S[i + 1] = S_new
    return err, info


def run_device(n_sites, J, h_fields, tau, n_periods, epsilon=0.0, state='neel', up_index=1,
               mode='reference', trunc=None, chi_cap=None, precondition=True):
    """The reference's Floquet sequence with the device update rule; also records the sweep counts."""
    kick, gates = tebd_ref.make_gates(n_sites, J, h_fields, tau, epsilon)
    psi0 = tebd_ref.product_state(n_sites, state, up_index)
    psi = psi0.copy()
    Z, Sent, LE, chi, sweeps = [], [], [], [], []
    L = n_sites
    for t in range(n_periods + 1):
        if t > 0:
            for half in range(2):
                for start in (0, 1):
                    for i in range(start, L - 1, 2):
                        _, info = update_bond_device(psi, i, gates[i], mode, trunc, chi_cap, precondition)
                        sweeps.append(info['sweeps'])
                if half == 0:
                    for i in range(L):
                        psi.apply_local_op(i, kick, unitary=True)
        Z.append(tebd_ref.site_z(psi))
        Sent.append(psi.entanglement_entropy())
        LE.append(abs(psi0.overlap(psi)) ** 2)
        chi.append(list(psi.chi))
    return dict(Z=np.array(Z), S_ent=np.array(Sent), LE=np.array(LE),
                chi=np.array(chi, dtype=int).reshape(len(chi), -1), psi=psi, sweeps=np.array(sweeps))
