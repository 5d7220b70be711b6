"""
NumPy model of the DEVICE algorithm, stage by stage -- TEST INFRASTRUCTURE.

The CUDA path does not call LAPACK.  Per two-site update it runs

  K1  C      = gate . (B_i B_{i+1})                      (no left Schmidt values)
  K2a G      = C^H diag(S_l^2) C                         (Gram matrix of theta = S_l C)
  K2b G      = Q T Q^H   Householder tridiagonalisation  (zhetd2 'L' convention)
      Q      generated backward from the reflectors      (zung2r convention)
  K2c T      = Z diag(lam) Z^T  implicit QL with Wilkinson shift (tql2)
      sort lam descending, truncate, S = sqrt(lam)/norm
  K2d B_{i+1}= (Q Z[:, kept])^H
  K3  B_i    = C . B_{i+1}^H / renorm                    (inverse-free Hastings update)

This file restates each stage in plain NumPy loops so the GPU kernel tests can
compare intermediate buffers (d, e, tau, Q, lam, ...) stage by stage, and so the
accuracy of the Gram route versus LAPACK zgesdd can be studied on the CPU.
"""

import numpy as np

from . import tebd_ref

EPS = np.finfo(float).eps


def larfg(alpha, x):
    """LAPACK zlarfg: returns (beta, tau, v_tail) with H = I - tau [1;v][1;v]^H,
    H^H [alpha; x] = [beta; 0], beta real."""
    xnorm = np.linalg.norm(x)
    if xnorm == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0 + 0.0j, np.zeros_like(x)
    beta = -np.copysign(np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm ** 2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    v = x / (alpha - beta)
    return beta, tau, v


def hetrd_lower(G):
    """Unblocked Hermitian tridiagonalisation (zhetd2, UPLO='L').
    Returns d[n], e[n-1], tau[n-1], V (n x n, column k holds reflector k in rows k+1..n-1, v[k+1]=1)."""
    A = np.array(G, dtype=complex)
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(max(n - 1, 0), dtype=complex)
    V = np.zeros((n, n), dtype=complex)
    for k in range(n - 1):
        alpha = A[k + 1, k]
        beta, t, vt = larfg(alpha, A[k + 2:, k])
        e[k] = beta
        tau[k] = t
        v = np.concatenate(([1.0 + 0j], vt))
        V[k + 1:, k] = v
        d[k] = A[k, k].real
        if t != 0:
            A22 = A[k + 1:, k + 1:]
            p = t * (A22 @ v)
            alpha2 = -0.5 * t * np.vdot(p, v)          # zdotc(p, v) = p^H v
            w = p + alpha2 * v
            A22 -= np.outer(v, w.conj()) + np.outer(w, v.conj())
    d[n - 1] = A[n - 1, n - 1].real
    return d, e, tau, V


def ungtr_lower(tau, V):
    """Q = H(0) H(1) ... H(n-2), generated backward (zung2r on the trailing block)."""
    n = V.shape[0]
    Q = np.eye(n, dtype=complex)
    for k in range(n - 2, -1, -1):
        if tau[k] == 0:
            continue
        v = V[k + 1:, k]
        blk = Q[k + 1:, k + 1:]
        t = v.conj() @ blk
        blk -= tau[k] * np.outer(v, t)
    return Q


def tql_implicit(d, e, max_iter=60):
    """Implicit QL with Wilkinson shift on a real symmetric tridiagonal (EISPACK tql2).
    Returns (lam unsorted, Z) with T = Z diag(lam) Z^T."""
    d = np.array(d, dtype=float)
    n = len(d)
    ee = np.zeros(n)
    ee[:n - 1] = e
    e = ee
    Z = np.eye(n)
    for l in range(n):
        it = 0
        while True:
            m = l
            while m < n - 1:
                dd = abs(d[m]) + abs(d[m + 1])
                if abs(e[m]) <= EPS * dd:
                    break
                m += 1
            if m == l:
                break
            it += 1
            if it > max_iter:
                raise RuntimeError("tql: no convergence")
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = np.hypot(g, 1.0)
            g = d[m] - d[l] + e[l] / (g + np.copysign(r, g))
            s = c = 1.0
            p = 0.0
            i = m - 1
            underflow = False
            while i >= l:
                f = s * e[i]
                b = c * e[i]
                r = np.hypot(f, g)
                e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= p
                    e[m] = 0.0
                    underflow = True
                    break
                s = f / r
                c = g / r
                g = d[i + 1] - p
                r = (d[i] - g) * s + 2.0 * c * b
                p = s * r
                d[i + 1] = g + p
                g = c * r - b
                zi1 = Z[:, i + 1].copy()
                Z[:, i + 1] = s * Z[:, i] + c * zi1
                Z[:, i] = c * Z[:, i] - s * zi1
                i -= 1
            if underflow:
                continue
            d[l] -= p
            e[l] = g
            e[m] = 0.0
    return d, Z


def truncate_device(lam_sorted, mode, cutoff=1e-13, chi_max=None, svd_min=None, trunc_cut=None,
                    chi_cap=None, gram_floor=1e-7):
    """Device truncation rule on eigenvalues sorted descending (lam = sigma^2).
    Returns (k, S_new, renorm, trunc_err)."""
    lam = np.maximum(np.asarray(lam_sorted, dtype=float), 0.0)
    sig = np.sqrt(lam)
    tot = np.sqrt(np.sum(lam))
    n = len(sig)
    floor = gram_floor * sig[0]
    if mode == 'reference':
        thr = max(cutoff, floor)
        k = int(np.sum(sig > thr))
        k = max(k, 1)
        if chi_cap is not None:
            k = min(k, chi_cap)
    else:
        # TeNPy truncate() on the normalised spectrum, written for a descending array: every
        # constraint is a maximal keep-count; a constraint that would keep nothing is ignored.
        s = sig / tot
        k = n if chi_max is None else min(n, chi_max)
        if chi_cap is not None:
            k = min(k, chi_cap)
        smin = max(svd_min or 0.0, floor / tot)
        k_svd = int(np.sum(s > smin))
        if k_svd >= 1:
            k = min(k, k_svd)
        if trunc_cut is not None:
            c = np.cumsum(s[::-1] ** 2)                  # ascending cumulative weight
            good = c > trunc_cut * trunc_cut
            if np.any(good):
                k = min(k, n - int(np.argmax(good)))
        k = max(k, 1)
    kept = sig[:k]
    renorm = np.linalg.norm(kept)
    err = float(np.sum(lam[k:]) / max(np.sum(lam), 1e-300))
    return k, kept / renorm, renorm, err


def gram_eig(C, S_left):
    """Stages K2a-K2c: returns (lam desc, V) with V the eigenvectors of G."""
    w = np.repeat(np.asarray(S_left, dtype=float) ** 2, C.shape[0] // len(S_left))
    G = (C.conj().T * w) @ C
    d, e, tau, Vr = hetrd_lower(G)
    Q = ungtr_lower(tau, Vr)
    lam, Z = tql_implicit(d, e)
    order = np.argsort(-lam, kind='stable')
    return lam[order], Q @ Z[:, order], dict(G=G, d=d, e=e, tau=tau, Q=Q, lam_unsorted=lam, Z=Z, order=order)


def update_bond_device(psi, i, gate, mode='reference', trunc=None, fast_eigh=False, chi_cap=None,
                       gram_floor=1e-7):
    """One two-site update with the device algorithm on an all-'B'-form oracle MPS."""
    trunc = trunc or {}
    B0 = psi.get_B(i, 'B')
    B1 = psi.get_B(i + 1, 'B')
    chiL, chiR = B0.shape[0], B1.shape[2]
    C = np.tensordot(B0, B1, axes=(2, 0))
    C = np.einsum('pqrs,arsb->apqb', np.asarray(gate).reshape(2, 2, 2, 2), C).reshape(2 * chiL, 2 * chiR)
    if fast_eigh:
        w = np.repeat(psi._S[i] ** 2, 2)
        G = (C.conj().T * w) @ C
        lam, V = np.linalg.eigh(G)
        lam, V = lam[::-1], V[:, ::-1]
    else:
        lam, V, _ = gram_eig(C, psi._S[i])
    k, S_new, renorm, err = truncate_device(lam, mode, chi_cap=chi_cap, gram_floor=gram_floor, **trunc)
    Vk = V[:, :k]
    psi._B[i + 1] = Vk.conj().T.reshape(k, 2, chiR)
    psi._B[i] = (C @ Vk / renorm).reshape(chiL, 2, k)
    psi.form[i] = psi.form[i + 1] = (0.0, 1.0)
    psi._S[i + 1] = S_new
    return err


def floquet_step_device(psi, kick, gates, mode='reference', trunc=None, fast_eigh=True, chi_cap=None,
                        gram_floor=1e-7):
    psi = psi.copy()
    L = psi.L
    for half in range(2):
        for start in (0, 1):
            for i in range(start, L - 1, 2):
                update_bond_device(psi, i, gates[i], mode, trunc, fast_eigh, chi_cap, gram_floor)
        if half == 0:
            for i in range(L):
                psi.apply_local_op(i, kick, unitary=True)
    return psi


def run_device(n_sites, J, h_fields, tau, n_periods, epsilon=0.0, state='neel', up_index=1,
               mode='reference', trunc=None, fast_eigh=True, chi_cap=None, gram_floor=1e-7):
    kick, gates = tebd_ref.make_gates(n_sites, J, h_fields, tau, epsilon)
    psi0 = tebd_ref.product_state(n_sites, state, up_index)
    psi = psi0.copy()
    Z, Sent, LE, chi = [], [], [], []
    for t in range(n_periods + 1):
        if t > 0:
            psi = floquet_step_device(psi, kick, gates, mode, trunc, fast_eigh, chi_cap, gram_floor)
        Z.append(tebd_ref.site_z(psi))
        Sent.append(psi.entanglement_entropy())
        LE.append(abs(psi0.overlap(psi)) ** 2)
        chi.append(list(psi.chi))
    return dict(Z=np.array(Z), S_ent=np.array(Sent), LE=np.array(LE),
                chi=np.array(chi, dtype=int).reshape(len(chi), -1), psi=psi)


# ===========================================================================
# Route D (the one the device uses): Golub-Kahan-Reinsch SVD, right vectors only.
#
#   K2a  theta = Qb B P^H     Householder bidiagonalisation (zgebd2, m >= n; theta is
#                             zero-padded to n rows when m < n); only d, e, the right
#                             reflectors and taup are kept
#   K2b  P generated backward from the right reflectors (zungbr 'P' content)
#   K2c  B = U_B diag(sig) Z^T   implicit-shift QR on the real bidiagonal (Golub-Reinsch
#                             as in EISPACK svd / "svdcmp"), accumulating ONLY the right
#                             rotations into Z
#   K2d  B_{i+1} = (P Z[:, kept])^H
#
# The Gram route above (kept for the record) loses a factor 1/sigma_k of accuracy in the
# vectors of small kept singular values and drifts from the LAPACK oracle by ~1e-7 over
# 60 periods at chi_max=32 (see tests/test_oracle_cpu.py), hence route D.
# ===========================================================================

def gebd2_right(theta):
    """Unblocked complex bidiagonalisation (zgebd2, upper bidiagonal, m >= n after padding).
    Returns d[n], e[n-1] (real), taup[n-1], Ur (n x n; row i holds right reflector i in
    columns i+1..n-1 with Ur[i,i+1] = 1)."""
    A = np.array(theta, dtype=complex)
    m, n = A.shape
    if m < n:
        A = np.vstack([A, np.zeros((n - m, n), dtype=complex)])
        m = n
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    taup = np.zeros(max(n - 1, 0), dtype=complex)
    Ur = np.zeros((n, n), dtype=complex)
    for i in range(n):
        # left reflector H(i) annihilates A[i+1:, i]
        beta, tq, vt = larfg(A[i, i], A[i + 1:, i])
        d[i] = beta
        if i < n - 1:
            v = np.concatenate(([1.0 + 0j], vt))
            if tq != 0:
                blk = A[i:, i + 1:]
                y = v.conj() @ blk                       # v^H A
                blk -= np.conj(tq) * np.outer(v, y)      # H^H applied from the left
            # right reflector G(i) annihilates A[i, i+2:]
            row = A[i, i + 1:].conj()
            beta, tp, ut = larfg(row[0], row[1:])
            e[i] = beta
            taup[i] = tp
            u = np.concatenate(([1.0 + 0j], ut))
            Ur[i, i + 1:] = u
            if tp != 0:
                blk = A[i + 1:, i + 1:]
                x = blk @ u
                blk -= tp * np.outer(x, u.conj())
    return d, e, taup, Ur


def ungbr_p(taup, Ur):
    """P = G(0) G(1) ... G(n-2), G(i) = I - taup_i u_i u_i^H, generated backward."""
    n = Ur.shape[0]
    P = np.eye(n, dtype=complex)
    for i in range(n - 2, -1, -1):
        if taup[i] == 0:
            continue
        u = Ur[i, i + 1:]
        blk = P[i + 1:, i + 1:]
        t = u.conj() @ blk
        blk -= taup[i] * np.outer(u, t)
    return P


def _pythag(a, b):
    return np.hypot(a, b)


def bdsqr_right(d, e, max_iter=75):
    """Golub-Reinsch implicit-shift QR on the upper bidiagonal (d, e), right rotations only.
    Returns (w, Z): singular values (unsorted, >= 0) and Z with B^T B = Z diag(w^2) Z^T."""
    n = len(d)
    w = np.array(d, dtype=float)
    rv1 = np.zeros(n)
    rv1[1:] = e                                  # rv1[i] couples w[i-1], w[i]
    Z = np.eye(n)
    anorm = 0.0
    for i in range(n):
        anorm = max(anorm, abs(w[i]) + abs(rv1[i]))
    for k in range(n - 1, -1, -1):
        for its in range(max_iter):
            flag = True
            l = k
            while True:
                nm = l - 1
                if abs(rv1[l]) + anorm == anorm:
                    flag = False
                    break
                if abs(w[nm]) + anorm == anorm:
                    break
                l -= 1
            if flag:
                # w[nm] negligible: cancel rv1[l] with left rotations (U not accumulated)
                c = 0.0
                s = 1.0
                for i in range(l, k + 1):
                    f = s * rv1[i]
                    rv1[i] = c * rv1[i]
                    if abs(f) + anorm == anorm:
                        break
                    g = w[i]
                    hh = _pythag(f, g)
                    w[i] = hh
                    hh = 1.0 / hh
                    c = g * hh
                    s = -f * hh
            z = w[k]
            if l == k:
                if z < 0.0:
                    w[k] = -z
                    Z[:, k] = -Z[:, k]
                break
            if its == max_iter - 1:
                raise RuntimeError("bdsqr: no convergence")
            x = w[l]
            nm = k - 1
            y = w[nm]
            g = rv1[nm]
            hh = rv1[k]
            f = ((y - z) * (y + z) + (g - hh) * (g + hh)) / (2.0 * hh * y)
            g = _pythag(f, 1.0)
            f = ((x - z) * (x + z) + hh * ((y / (f + np.copysign(g, f))) - hh)) / x
            c = s = 1.0
            for j in range(l, nm + 1):
                i = j + 1
                g = rv1[i]
                y = w[i]
                hh = s * g
                g = c * g
                z = _pythag(f, hh)
                rv1[j] = z
                c = f / z
                s = hh / z
                f = x * c + g * s
                g = g * c - x * s
                hh = y * s
                y *= c
                zj = Z[:, j].copy()
                zi = Z[:, i].copy()
                Z[:, j] = zj * c + zi * s
                Z[:, i] = zi * c - zj * s
                z = _pythag(f, hh)
                w[j] = z
                if z != 0.0:
                    z = 1.0 / z
                    c = f * z
                    s = hh * z
                f = c * g + s * y
                x = c * y - s * g
            rv1[l] = 0.0
            rv1[k] = f
            w[k] = x
    return w, Z


def svd_right(theta):
    """Route-D SVD: returns (sig desc, V) with theta^H theta = V diag(sig^2) V^H."""
    d, e, taup, Ur = gebd2_right(theta)
    P = ungbr_p(taup, Ur)
    w, Z = bdsqr_right(d, e)
    order = np.argsort(-w, kind='stable')
    return w[order], P @ Z[:, order], dict(d=d, e=e, taup=taup, Ur=Ur, P=P, w=w, Z=Z, order=order)


def update_bond_device_D(psi, i, gate, mode='reference', trunc=None, chi_cap=None, lapack=False):
    """Two-site update with route D (lapack=True swaps in scipy's SVD for the V factor:
    same truncation rule, used to separate rule effects from SVD-accuracy effects)."""
    trunc = trunc or {}
    B0 = psi.get_B(i, 'B')
    B1 = psi.get_B(i + 1, 'B')
    chiL, chiR = B0.shape[0], B1.shape[2]
    C = np.tensordot(B0, B1, axes=(2, 0))
    C = np.einsum('pqrs,arsb->apqb', np.asarray(gate).reshape(2, 2, 2, 2), C).reshape(2 * chiL, 2 * chiR)
    theta = C * np.repeat(psi._S[i], 2)[:, None]
    if lapack:
        _, sig, Vh = tebd_ref._svd(theta)
        V = Vh.conj().T
        if len(sig) < theta.shape[1]:
            sig = np.concatenate([sig, np.zeros(theta.shape[1] - len(sig))])
            V = np.hstack([V, np.zeros((V.shape[0], theta.shape[1] - V.shape[1]))])
    else:
        sig, V, _ = svd_right(theta)
    k, S_new, renorm, err = truncate_device(sig ** 2, mode, chi_cap=chi_cap, gram_floor=0.0, **trunc)
    Vk = V[:, :k]
    psi._B[i + 1] = Vk.conj().T.reshape(k, 2, chiR)
    psi._B[i] = (C @ Vk / renorm).reshape(chiL, 2, k)
    psi.form[i] = psi.form[i + 1] = (0.0, 1.0)
    psi._S[i + 1] = S_new
    return err


def run_device_D(n_sites, J, h_fields, tau, n_periods, epsilon=0.0, state='neel', up_index=1,
                 mode='reference', trunc=None, chi_cap=None, lapack=False):
    kick, gates = tebd_ref.make_gates(n_sites, J, h_fields, tau, epsilon)
    psi0 = tebd_ref.product_state(n_sites, state, up_index)
    psi = psi0.copy()
    Z, Sent, LE, chi = [], [], [], []
    L = n_sites
    for t in range(n_periods + 1):
        if t > 0:
            for half in range(2):
                for start in (0, 1):
                    for i in range(start, L - 1, 2):
                        update_bond_device_D(psi, i, gates[i], mode, trunc, chi_cap, lapack)
                if half == 0:
                    for i in range(L):
                        psi.apply_local_op(i, kick, unitary=True)
        Z.append(tebd_ref.site_z(psi))
        Sent.append(psi.entanglement_entropy())
        LE.append(abs(psi0.overlap(psi)) ** 2)
        chi.append(list(psi.chi))
    return dict(Z=np.array(Z), S_ent=np.array(Sent), LE=np.array(LE),
                chi=np.array(chi, dtype=int).reshape(len(chi), -1), psi=psi)
