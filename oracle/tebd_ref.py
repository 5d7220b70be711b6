"""
O2 -- NumPy/LAPACK restatement of the MPS runtime the reference sits on.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity unpinned at the TeNPy
boundary: TeNPy (physics-tenpy v1.0.x) is a third-party dependency that is
absent from /root/reference and from this image; what follows restates its
published algorithms as used by the reference's call sites:

  src/models/kicked_ising.py:186,206   MPS.apply_local_op(i, op, unitary=True)
  src/core/tensor_utils.py:60          MPS.from_product_state
  src/core/observables.py:25           MPS.overlap
  src/core/observables.py:62,68        MPS.expectation_value(op, sites=[i])
  src/core/observables.py:121          MPS.correlation_function
  src/core/observables.py:250          MPS.get_SL
  src/core/tensor_utils.py:180         MPS.entanglement_entropy

Two update rules are provided:

* ``apply_local_op``      the reference's live path:
                          theta with both outer S, LAPACK SVD, absolute cutoff
                          1e-13, S renormalised, mixed 'A'/'B' forms with lazy
                          S**-1 conversion.
* ``update_bond_tebd``    TeNPy's ``TEBDEngine.update_bond``:
                          inverse-free B-form update with ``truncate()``
                          semantics for chi_max / svd_min / trunc_cut.  This is
                          what "equal chi_max and truncation cutoff" means.
"""

import numpy as np
import scipy.linalg

_FORMS = {'A': (1.0, 0.0), 'B': (0.0, 1.0), 'C': (0.5, 0.5), 'G': (0.0, 0.0),
          'Th': (1.0, 1.0), None: None}


def _svd(a):
    """LAPACK zgesdd with zgesvd fallback (what npc.svd does [TeNPy-mem])."""
    try:
        return scipy.linalg.svd(a, full_matrices=False, lapack_driver='gesdd')
    except np.linalg.LinAlgError:
        return scipy.linalg.svd(a, full_matrices=False, lapack_driver='gesvd')


def truncate(S, chi_max=None, svd_min=None, trunc_cut=None, chi_min=None):
    """TeNPy ``tenpy.algorithms.truncation.truncate`` [TeNPy-mem].

    ``S`` must be normalised (sum S**2 == 1).  Returns (mask, new_norm, err)
    with ``err`` the discarded weight sum(S_discarded**2).
    """
    S = np.asarray(S, dtype=float)
    if trunc_cut is not None and trunc_cut >= 1.0:
        raise ValueError("trunc_cut >= 1")
    with np.errstate(divide='ignore'):
        logS = np.log(S)
    piv = np.argsort(logS)          # ascending: smallest first
    logS = logS[piv]
    good = np.ones(len(piv), dtype=bool)

    def combine(g1, g2):
        res = np.logical_and(g1, g2)
        return res if np.any(res) else g1

    if chi_max is not None:
        g2 = np.zeros(len(piv), dtype=bool)
        g2[-chi_max:] = True
        good = combine(good, g2)
    if chi_min is not None and chi_min > 1:
        g2 = np.ones(len(piv), dtype=bool)
        g2[-chi_min + 1:] = False
        good = combine(good, g2)
    if svd_min is not None:
        g2 = np.exp(logS) > svd_min
        good = combine(good, g2)
    if trunc_cut is not None:
        g2 = np.cumsum(S[piv] ** 2) > trunc_cut * trunc_cut
        good = combine(good, g2)
    cut = np.nonzero(good)[0][0]
    mask = np.zeros(len(S), dtype=bool)
    mask[piv[cut:]] = True
    new_norm = np.linalg.norm(S[mask])
    err = float(np.sum(S[~mask] ** 2))
    return mask, new_norm, err


class MPS:
    """Finite MPS with TeNPy's storage convention: site tensors (vL, p, vR),
    Schmidt values S[0..L], per-site canonical form exponents (nuL, nuR) such
    that stored = S_left**nuL * Gamma * S_right**nuR."""

    def __init__(self, sites, Bs, Ss, forms, norm=1.0):
        self.sites = list(sites)
        self._B = [np.array(b) for b in Bs]
        self._S = [np.array(s, dtype=float) for s in Ss]
        self.form = [tuple(f) for f in forms]
        self.norm = norm
        self.bc = 'finite'
        self.dtype = np.result_type(*[b.dtype for b in self._B]) if self._B else np.float64

    # ------------------------------------------------------------------ basics
    @property
    def L(self):
        return len(self._B)

    @property
    def chi(self):
        return [int(len(s)) for s in self._S[1:-1]]

    def copy(self):
        return MPS(self.sites, [b.copy() for b in self._B], [s.copy() for s in self._S],
                   list(self.form), self.norm)

    @classmethod
    def from_product_state(cls, sites, p_state, bc='finite', dtype=np.float64):
        """L tensors (1,d,1), one-hot; S = [1.] everywhere; all 'B' form (A.2.1)."""
        Bs = []
        for site, st in zip(sites, p_state):
            d = getattr(site, 'dim', 2)
            idx = site.state_index(st) if hasattr(site, 'state_index') else int(st)
            b = np.zeros((1, d, 1), dtype=dtype)
            b[0, idx, 0] = 1.0
            Bs.append(b)
        L = len(Bs)
        return cls(sites, Bs, [np.ones(1)] * (L + 1), [_FORMS['B']] * L)

    # ------------------------------------------------------- form conversions
    def get_SL(self, i):
        return self._S[i]

    def get_SR(self, i):
        return self._S[i + 1]

    def get_B(self, i, form='B'):
        """Tensor of site i converted to `form` (tuple, label, or None = as stored).
        A tuple entry None keeps that side as stored."""
        if isinstance(form, str) or form is None:
            form = _FORMS[form]
        B = self._B[i]
        if form is None:
            return B
        oL, oR = self.form[i]
        nL = oL if form[0] is None else form[0]
        nR = oR if form[1] is None else form[1]
        if nL != oL:
            B = self._scale(B, self._S[i], nL - oL, 0)
        if nR != oR:
            B = self._scale(B, self._S[i + 1], nR - oR, 2)
        return B

    @staticmethod
    def _scale(B, S, diff, axis):
        if diff == -1.0:
            f = 1.0 / S
        elif diff == 1.0:
            f = S
        else:
            f = S ** diff
        shape = [1, 1, 1]
        shape[axis] = len(S)
        return B * f.reshape(shape)

    def get_theta(self, i, n=2, formL=1.0, formR=1.0):
        """S_i**formL B_i ... B_{i+n-1} S_{i+n}**formR with inner S exactly once (A.2.3)."""
        if n == 1:
            return self.get_B(i, (formL, formR))
        theta = self.get_B(i, (formL, None))
        old_fR = self.form[i][1]
        for k in range(1, n):
            j = i + k
            new_fR = None if k + 1 < n else formR
            B = self.get_B(j, (1.0 - old_fR, new_fR))
            old_fR = self.form[j][1]
            theta = np.tensordot(theta, B, axes=(theta.ndim - 1, 0))
        return theta            # (vL, p0, ..., p_{n-1}, vR)

    # -------------------------------------------------------------- local ops
    def apply_local_op(self, i, op, unitary=None, renormalize=False, cutoff=1.e-13):
        """The reference's gate application (kicked_ising.py:186,206)."""
        if i < 0:
            i += self.L
        if not (0 <= i < self.L):
            raise IndexError("site index out of range")
        op = np.asarray(op.to_ndarray() if hasattr(op, 'to_ndarray') else op)
        n = op.ndim // 2
        if i + n > self.L:
            raise ValueError("local operator does not fit on finite MPS")
        if n == 1:
            # B[vL,p,vR] <- sum_p' op[p,p'] B[vL,p',vR]; form label unchanged (A.2.2)
            self._B[i] = np.einsum('pq,aqb->apb', op, self._B[i])
        elif n == 2:
            th = self.get_theta(i, 2)                       # (vL,p0,p1,vR), both outer S
            th = np.einsum('pqrs,arsb->apqb', op, th)
            chiL, d0, d1, chiR = th.shape
            U, S, Vh = _svd(th.reshape(chiL * d0, d1 * chiR))
            keep = S > cutoff                               # absolute cutoff
            U, S, Vh = U[:, keep], S[keep], Vh[keep, :]
            S = S / np.linalg.norm(S)                       # from_full: S /= norm(S)
            k = len(S)
            self._B[i] = U.reshape(chiL, d0, k)
            self.form[i] = _FORMS['A']
            self._B[i + 1] = Vh.reshape(k, d1, chiR)
            self.form[i + 1] = _FORMS['B']
            self._S[i + 1] = S
        else:
            raise NotImplementedError("only 1- and 2-site operators")
        if unitary is None:
            unitary = False
        if not unitary:
            raise NotImplementedError("non-unitary apply_local_op (canonical_form) not restated")
        self.dtype = np.result_type(self.dtype, op.dtype)

    def update_bond_tebd(self, i, gate, chi_max=None, svd_min=None, trunc_cut=None):
        """TeNPy TEBDEngine.update_bond on sites (i, i+1): inverse-free, all-B form
        (A.2.4) with truncate() (A.2.5).  Returns the truncation error (weight)."""
        gate = np.asarray(gate).reshape(2, 2, 2, 2)
        B0 = self.get_B(i, 'B')
        B1 = self.get_B(i + 1, 'B')
        C = np.tensordot(B0, B1, axes=(2, 0))               # (vL,p0,p1,vR)
        C = np.einsum('pqrs,arsb->apqb', gate, C)
        theta = C * self._S[i].reshape(-1, 1, 1, 1)
        chiL, d0, d1, chiR = theta.shape
        U, S, Vh = _svd(theta.reshape(chiL * d0, d1 * chiR))
        renorm = np.linalg.norm(S)
        S = S / renorm
        mask, new_norm, err = truncate(S, chi_max=chi_max, svd_min=svd_min, trunc_cut=trunc_cut)
        S = S[mask] / new_norm
        renorm *= new_norm
        Vh = Vh[mask, :]
        k = len(S)
        BL = np.tensordot(C.reshape(chiL, d0, d1 * chiR), Vh.conj(), axes=(2, 1)) / renorm
        self._B[i] = BL
        self.form[i] = _FORMS['B']
        self._B[i + 1] = Vh.reshape(k, d1, chiR)
        self.form[i + 1] = _FORMS['B']
        self._S[i + 1] = S
        return err

    # ------------------------------------------------------------ observables
    def overlap(self, other):
        """<self|other> by left-to-right transfer contraction (A.2.6)."""
        if self.L != other.L:
            raise ValueError("length mismatch")
        E = np.ones((1, 1), dtype=complex)
        for i in range(self.L):
            # state = S0 G0 S1 G1 S2 ... : site 0 in (1,1) form, the rest in 'B' form
            a = self.get_B(i, 'B') if i > 0 else self.get_B(0, (1.0, 1.0))
            b = other.get_B(i, 'B') if i > 0 else other.get_B(0, (1.0, 1.0))
            T = np.tensordot(E, b, axes=(1, 0))             # (a, p, b')
            E = np.tensordot(a.conj(), T, axes=((0, 1), (0, 1)))
        return complex(E[0, 0]) * self.norm * other.norm

    def expectation_value(self, ops, sites=None):
        """<theta_i|op|theta_i>, theta_i = S_i B_i (A.2.7).  Single-site ops only."""
        if sites is None:
            sites = range(self.L)
        out = []
        for i in sites:
            op = ops
            if isinstance(op, str):
                op = self.sites[i].get_op(op)
            op = np.asarray(op.to_ndarray() if hasattr(op, 'to_ndarray') else op)
            th = self.get_theta(i, 1)
            val = np.einsum('apb,pq,aqb->', th.conj(), op, th)
            out.append(val)
        out = np.array(out)
        return np.real_if_close(out)

    def correlation_function(self, op1, op2, sites1=None, sites2=None):
        """<op1_i op2_j> for i in sites1, j in sites2 (transfer contraction)."""
        if sites1 is None:
            sites1 = range(self.L)
        if sites2 is None:
            sites2 = range(self.L)
        o1 = np.asarray(op1.to_ndarray() if hasattr(op1, 'to_ndarray') else op1)
        o2 = np.asarray(op2.to_ndarray() if hasattr(op2, 'to_ndarray') else op2)
        res = np.zeros((len(list(sites1)), len(list(sites2))), dtype=complex)
        for a, i in enumerate(sites1):
            for b, j in enumerate(sites2):
                res[a, b] = self._corr(o1, o2, i, j)
        return np.real_if_close(res)

    def _corr(self, o1, o2, i, j):
        if i == j:
            th = self.get_theta(i, 1)
            return np.einsum('apb,pq,aqb->', th.conj(), o1 @ o2, th)
        if i > j:
            i, j, o1, o2 = j, i, o2, o1
        th = self.get_B(i, (1.0, 0.0))                      # 'A' form incl. left S
        E = np.einsum('apb,pq,aqc->bc', th.conj(), o1, th)
        for k in range(i + 1, j):
            A = self.get_B(k, (1.0, 0.0))
            E = np.einsum('bc,bpd,cpe->de', E, A.conj(), A)
        th = self.get_B(j, (1.0, 1.0))
        return np.einsum('bc,bpd,pq,cqd->', E, th.conj(), o2, th)

    def entanglement_entropy(self):
        """von Neumann entropies of bonds 1..L-1, natural log, p>1e-30 (A.2.8)."""
        out = []
        for S in self._S[1:-1]:
            p = S ** 2
            p = p[p > 1.e-30]
            out.append(-np.sum(p * np.log(p)))
        return np.array(out)

    # --------------------------------------------------------------- helpers
    def to_statevector(self):
        """Dense amplitudes psi[p0,...,p_{L-1}] flattened (for cross-checks, small L)."""
        v = self.get_B(0, (1.0, 1.0))
        for i in range(1, self.L):
            v = np.tensordot(v, self.get_B(i, 'B'), axes=(v.ndim - 1, 0))
        return v.reshape(-1) * self.norm


# --------------------------------------------------------------------------
# The reference's Floquet sequence restated on top of the MPS above.
# --------------------------------------------------------------------------

SIGMA_X = np.array([[0, 1], [1, 0]], dtype=complex)
SIGMA_Z = np.array([[1, 0], [0, -1]], dtype=complex)
SIGMA_I = np.eye(2, dtype=complex)


def disorder_fields(n_sites, h_disorder, seed):
    """kicked_ising.py:55-59 -- legacy global NumPy RNG, reseeded."""
    if seed is not None:
        np.random.seed(seed)
    return np.random.uniform(-h_disorder, h_disorder, n_sites)


def make_gates(n_sites, J, h_fields, tau, epsilon=0.0):
    """kicked_ising.py:73-90.  epsilon != 0 gives the imperfect pulse
    expm(-i (pi/2)(1-eps) sigma_x) (not in the reference)."""
    kick = scipy.linalg.expm(-1j * np.pi / 2 * (1.0 - epsilon) * SIGMA_X)
    gates = []
    for i in range(n_sites - 1):
        h2 = (J * np.kron(SIGMA_Z, SIGMA_Z) + h_fields[i] * np.kron(SIGMA_Z, SIGMA_I)
              + h_fields[i + 1] * np.kron(SIGMA_I, SIGMA_Z))
        gates.append(scipy.linalg.expm(-1j * tau / 2 * h2))
    return kick, gates


def product_state(n_sites, state_type='neel', up_index=1, rng_choice=None):
    """tensor_utils.py:44-60.  `up_index` = internal basis index TeNPy assigns to
    the label 'up' for SpinHalfSite(conserve='parity')."""
    if state_type == 'all_up':
        labels = ['up'] * n_sites
    elif state_type == 'all_down':
        labels = ['down'] * n_sites
    elif state_type == 'neel':
        labels = ['up' if i % 2 == 0 else 'down' for i in range(n_sites)]
    elif state_type == 'random':
        labels = [np.random.choice(['up', 'down']) for _ in range(n_sites)]
    else:
        raise ValueError(f"Unknown state type: {state_type}")
    idx = [up_index if s == 'up' else 1 - up_index for s in labels]
    sites = [None] * n_sites
    return MPS.from_product_state(sites, idx)


def floquet_step(psi, kick, gates, mode='reference', trunc=None):
    """One Floquet period, same order as kicked_ising.py:100-160:
    even bonds, odd bonds, kick on every site, even bonds, odd bonds.
    mode='reference': apply_local_op (cutoff 1e-13, trunc ignored);
    mode='tebd'     : update_bond_tebd with trunc = dict(chi_max, svd_min, trunc_cut).
    Returns (new_psi, truncation_error_sum)."""
    psi = psi.copy()
    L = psi.L
    err = 0.0
    trunc = trunc or {}

    def two_site(i):
        nonlocal err
        if mode == 'reference':
            psi.apply_local_op(i, gates[i].reshape(2, 2, 2, 2), unitary=True)
        else:
            err += psi.update_bond_tebd(i, gates[i], chi_max=trunc.get('chi_max'),
                                        svd_min=trunc.get('svd_min'),
                                        trunc_cut=trunc.get('trunc_cut'))

    def ising():
        for i in range(0, L - 1, 2):
            two_site(i)
        for i in range(1, L - 1, 2):
            two_site(i)

    ising()
    for i in range(L):
        psi.apply_local_op(i, kick, unitary=True)
    ising()
    return psi, err


def site_z(psi):
    """<Z_i> for all sites with the raw diag(+1,-1) on the internal index (observables.py:41-62)."""
    return np.array([float(np.real(psi.expectation_value(SIGMA_Z, sites=[i])[0])) for i in range(psi.L)])


def run(n_sites, J, h_fields, tau, n_periods, epsilon=0.0, state='neel', up_index=1,
        mode='reference', trunc=None, measure_every=1):
    """Evolve and record the observables the north star names.  Returns dict with
    Z[t,i], S_ent[t,b] (bonds 1..L-1), LE[t], chi[t,b], times[t], trunc_err."""
    kick, gates = make_gates(n_sites, J, h_fields, tau, epsilon)
    psi0 = product_state(n_sites, state, up_index)
    psi = psi0.copy()
    Z, Sent, LE, chi, times = [], [], [], [], []
    terr = 0.0

    def measure(t):
        Z.append(site_z(psi))
        Sent.append(psi.entanglement_entropy())
        LE.append(abs(psi0.overlap(psi)) ** 2)
        chi.append(list(psi.chi))
        times.append(t * 2 * tau)

    measure(0)
    for t in range(n_periods):
        psi, e = floquet_step(psi, kick, gates, mode=mode, trunc=trunc)
        terr += e
        if t % measure_every == 0:
            measure(t + 1)
    return dict(Z=np.array(Z), S_ent=np.array(Sent), LE=np.array(LE),
                chi=np.array(chi, dtype=int).reshape(len(chi), -1), times=np.array(times),
                trunc_err=terr, psi=psi)


def run_schedule(n_sites, J, h_fields, tau, schedule, state='neel', up_index=1, mode='tebd', trunc=None):
    """`run` with a piecewise-constant kick imperfection: schedule = [(epsilon, n_periods), ...] (bench.py's
    two-phase workload: entangle at a large epsilon until the bonds saturate, then evolve at the target epsilon).
    A record before the first period and after every period; the Ising gates do not depend on epsilon."""
    _, gates = make_gates(n_sites, J, h_fields, tau, 0.0)
    psi0 = product_state(n_sites, state, up_index)
    psi = psi0.copy()
    Z, Sent, LE, chi, eps_t = [], [], [], [], []
    terr = 0.0

    def measure():
        Z.append(site_z(psi))
        Sent.append(psi.entanglement_entropy())
        LE.append(abs(psi0.overlap(psi)) ** 2)
        chi.append(list(psi.chi))

    measure()
    for eps, n in schedule:
        kick, _ = make_gates(2, J, np.zeros(2), tau, eps)
        for _ in range(n):
            psi, e = floquet_step(psi, kick, gates, mode=mode, trunc=trunc)
            terr += e
            eps_t.append(eps)
            measure()
    return dict(Z=np.array(Z), S_ent=np.array(Sent), LE=np.array(LE),
                chi=np.array(chi, dtype=int).reshape(len(chi), -1), eps=np.array(eps_t), trunc_err=terr, psi=psi)
