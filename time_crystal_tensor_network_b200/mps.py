"""MPS handle: the subset of TeNPy's ``MPS`` interface that the reference's modules and tests touch
(SURVEY 8a.18), backed by a single-chain engine context on the GPU.

Reference call sites this object answers: ``MPS.from_product_state`` (src/core/tensor_utils.py:60),
``copy`` (src/models/kicked_ising.py:115,...), ``apply_local_op`` (kicked_ising.py:186,206),
``overlap`` (src/core/observables.py:25), ``expectation_value`` (observables.py:62,68),
``correlation_function`` (observables.py:121), ``get_SL`` (observables.py:250),
``entanglement_entropy`` (src/core/tensor_utils.py:180), ``chi`` / ``L`` / ``norm`` / ``sites``.

All tensors live on the device in right-canonical form; nothing here computes on the host beyond
2x2 bookkeeping of results the kernels return.
"""
import os

import numpy as np

from .engine import Context, EngineError

#: internal basis index TeNPy >= 1.0 assigns to the label 'up' of SpinHalfSite(conserve='parity')
#: (sort_charge=True).  Set TC_UP_INDEX=0 for the pre-1.0 ordering.
UP_INDEX = int(os.environ.get('TC_UP_INDEX', 1))

#: hard ceiling for automatically grown bond dimensions (one context must hold (2 chi)^2 workspaces)
CHI_HARD_CAP = int(os.environ.get('TC_CHI_HARD_CAP', 1024))


class _Leg:
    def __init__(self, conserve, qconj=1):
        self.conserve, self.qconj = conserve, qconj

    def conj(self):
        return _Leg(self.conserve, -self.qconj)


class SpinHalfSite:
    """Stand-in for tenpy.networks.site.SpinHalfSite: label <-> internal-index bookkeeping only."""
    dim = 2

    def __init__(self, conserve='Sz', sort_charge=None):
        self.conserve = conserve
        # TeNPy >= 1.0 sorts the charges of both conserving variants (sort_charge=True), which puts 'up' at index 1;
        # without conservation the order stays ['up', 'down']
        self._up = UP_INDEX if conserve in ('parity', 'Sz') else 0
        self.leg = _Leg(conserve)
        self.state_labels = {'up': self._up, 'down': 1 - self._up}

    def state_index(self, label):
        return self.state_labels[label] if isinstance(label, str) else int(label)

    def get_op(self, name):
        sz = np.zeros((2, 2), dtype=complex)
        sz[self._up, self._up], sz[1 - self._up, 1 - self._up] = 0.5, -0.5
        sx = np.array([[0, 0.5], [0.5, 0]], dtype=complex)
        sy = np.zeros((2, 2), dtype=complex)
        sy[self._up, 1 - self._up], sy[1 - self._up, self._up] = -0.5j, 0.5j
        ops = {'Sz': sz, 'Sx': sx, 'Sy': sy, 'Sigmaz': 2 * sz, 'Sigmax': 2 * sx, 'Sigmay': 2 * sy,
               'Id': np.eye(2, dtype=complex)}
        return ops[name]


def _as_matrix(op, site=None):
    if isinstance(op, str):
        return (site or SpinHalfSite(None)).get_op(op)
    if hasattr(op, 'to_ndarray'):
        op = op.to_ndarray()
    return np.asarray(op, dtype=complex)


class MPS:
    """One finite chain on the GPU (a context with R = 1)."""

    bc = 'finite'

    def __init__(self, ctx, sites=None):
        self._ctx = ctx
        self.sites = list(sites) if sites is not None else [SpinHalfSite(None) for _ in range(ctx.L)]
        self.norm = 1.0
        self._cache = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_product_state(cls, sites, p_state, bc='finite', dtype=np.float64, device=0):
        sites = list(sites)
        if len(sites) != len(p_state):
            raise ValueError('need one state label per site')
        idx = [s.state_index(p) if hasattr(s, 'state_index') else int(p) for s, p in zip(sites, p_state)]
        if not idx:
            raise ValueError('an MPS needs at least one site')
        ctx = Context(len(idx), 1, 1, device)
        ctx.set_product_state(np.array(idx, dtype=np.int8)[None, :])
        return cls(ctx, sites)

    def copy(self, chi_cap=None, storage=False):
        """New, independent state with the same tensors (optionally with room for larger bonds).
        storage=True: a snapshot in a context without SVD workspace (tensors, Schmidt values and observable scratch
        only, sized to the current bond dimensions); it is measured, overlapped and copied like any other state and
        moves to a full context by itself the first time a gate is applied to it."""
        cap = max(int(chi_cap or 0), max(self._chi_full()), 1)
        ctx = Context(self.L, cap, 1, self._ctx.device, storage_only=storage)
        ctx.copy_chain_from(0, self._ctx, 0)
        out = MPS(ctx, self.sites)
        out.norm = self.norm
        return out

    def _grow(self, chi_cap):
        """Make sure the state sits in a full (evolvable) context that can hold bonds up to chi_cap."""
        if chi_cap > self._ctx.chi_cap or self._ctx.storage_only:
            ctx = Context(self.L, max(chi_cap, self._ctx.chi_cap), 1, self._ctx.device)
            ctx.copy_chain_from(0, self._ctx, 0)
            self._ctx.close()
            self._ctx = ctx

    def _touch(self):
        self._cache = {}

    # ------------------------------------------------------------------ structure
    @property
    def L(self):
        return self._ctx.L

    def _chi_full(self):
        if 'chi' not in self._cache:
            self._cache['chi'] = [int(x) for x in self._ctx.chi()[0]]
        return self._cache['chi']

    @property
    def chi(self):
        """Bond dimensions of the L-1 inner bonds (TeNPy convention; empty for L = 1)."""
        return self._chi_full()[1:-1]

    def get_B(self, i, form='B'):
        """Site tensor (chi_l, 2, chi_r) as complex128; 'B' (right-canonical, as stored), 'A', 'Th' or 'G'."""
        B = self._ctx.get_site(0, i)
        if form in ('B', None):
            return B
        sl, sr = self.get_SL(i), self.get_SR(i)
        if form == 'Th':
            return B * sl[:, None, None]
        if form == 'A':
            return B * sl[:, None, None] / sr[None, None, :]
        if form == 'G':
            return B / sr[None, None, :]
        raise ValueError(f'unknown canonical form {form!r}')

    def get_SL(self, i):
        """Schmidt values on the bond left of site i (observables.py:250)."""
        if i < 0:
            i += self.L
        return self._ctx.get_S(0, i)

    def get_SR(self, i):
        if i < 0:
            i += self.L
        return self._ctx.get_S(0, i + 1)

    # ------------------------------------------------------------------ gates
    def apply_local_op(self, i, op, unitary=None, renormalize=False, cutoff=1.e-13):
        """One- or two-site operator on sites i (, i+1) with TeNPy's apply_local_op semantics for the
        unitary case: two-site operators are followed by an SVD that keeps
        sigma > cutoff and renormalises."""
        if i < 0:
            i += self.L
        if not 0 <= i < self.L:
            raise IndexError('site index out of range')
        m = _as_matrix(op, self.sites[i])
        n = {4: 1, 16: 2}.get(m.size, 0)
        if not unitary:
            raise NotImplementedError('only unitary=True is supported (the reference never uses anything else)')
        if n == 1:
            self._ctx.apply_one_site(0, i, m.reshape(2, 2))
        elif n == 2:
            if i + 2 > self.L:
                raise ValueError('local operator does not fit on finite MPS')
            chi = self._chi_full()
            need = min(2 * chi[i], 2 * chi[i + 2])
            if need > CHI_HARD_CAP:
                raise EngineError(f'bond dimension {need} exceeds TC_CHI_HARD_CAP={CHI_HARD_CAP}')
            self._grow(max(need, max(chi)))
            self._ctx.set_trunc('reference', cutoff=cutoff)
            self._ctx.apply_two_site(0, i, m.reshape(4, 4))
        else:
            raise NotImplementedError('only 1- and 2-site operators')
        self._touch()

    def canonical_form(self, renormalize=True, tol=1e-13, max_sweeps=None):
        """Bring the chain back to canonical form after non-unitary gates (TeNPy: ``MPS.canonical_form``).

        The TEBD update of a bond assumes orthonormal surroundings; a non-unitary gate (imaginary time) breaks that for
        its neighbours at O(dt), and single-site expectation values / Schmidt values read off the tensors are then off
        by as much.  Sweeps of IDENTITY gates without truncation repair it with the kernels the evolution uses: the SVD
        of an identity update re-orthonormalises the two tensors of a bond and replaces its Schmidt values, the state
        itself (the product of the B tensors) does not change, and every even + odd sweep carries the orthonormality
        one bond further, so a finite chain is exactly canonical after at most L/2 + 1 sweeps and a short-range
        correlated one much earlier (NumPy model: tests/studies/recanonicalise_identity_layers.py).  The sweeps stop
        once no Schmidt value moves by more than ``tol``.  The model and the truncation rule of the context are
        overwritten (every evolution call sets its own).  The state comes back normalised (the kernels renormalise
        every update and do not keep the factor), hence ``renormalize=False`` is refused.  Returns the number of
        sweeps made."""
        if not renormalize:
            raise NotImplementedError('canonical_form keeps no norm factor: only renormalize=True')
        L = self.L
        if L < 2:
            return 0
        chi = self._chi_full()
        self._grow(max(chi))
        ctx = self._ctx
        ctx.set_model(np.eye(4, dtype=complex).reshape(1, 1, 4, 4).repeat(L - 1, axis=1),
                      np.eye(2, dtype=complex).reshape(1, 2, 2))
        # no bond can grow under identity gates; 1e-14 (relative) only drops the numerically zero directions
        ctx.set_trunc('tebd', chi_max=max(chi), svd_min=1e-14, trunc_cut=0.0)
        if max_sweeps is None:
            max_sweeps = L // 2 + 2
        schmidt = lambda: [ctx.get_S(0, b) for b in range(1, L)]
        old, sweeps = schmidt(), 0
        for sweeps in range(1, max_sweeps + 1):
            ctx.apply_layer(0, 0)
            if L > 2:
                ctx.apply_layer(1, 0)
            new = schmidt()
            moved = max((np.max(np.abs(a - b)) if a.shape == b.shape else np.inf) for a, b in zip(old, new))
            old = new
            if moved <= tol:
                break
        fl = ctx.flags()
        if fl['chi_cap_overflow'] or fl['svd_not_converged']:
            raise EngineError(f'canonical_form failed on the device: {fl}')
        self.norm = 1.0
        self._touch()
        return sweeps

    # ------------------------------------------------------------------ observables
    def _rdm(self):
        if 'rdm' not in self._cache:
            rdm, ent = self._ctx.measure(entropies=True)
            self._cache['rdm'] = rdm[0]
            self._cache['ent'] = ent[0] if ent is not None else np.zeros(0)
        return self._cache['rdm']

    def expectation_value(self, ops, sites=None):
        """<theta_i| op |theta_i> for single-site operators."""
        rdm = self._rdm()
        if sites is None:
            sites = range(self.L)
        out = []
        for i in sites:
            m = _as_matrix(ops, self.sites[i]).reshape(2, 2)
            r00, r11, re, im = rdm[i]
            # sum_{p,q} op[p,q] sum conj(theta_p) theta_q ; the kernel returns sum theta_0 conj(theta_1) = re + i im
            out.append(m[0, 0] * r00 + m[1, 1] * r11 + m[0, 1] * complex(re, -im) + m[1, 0] * complex(re, im))
        return np.real_if_close(np.array(out))

    def entanglement_entropy(self):
        """von Neumann entropies of the L-1 inner bonds, natural log."""
        self._rdm()
        return np.array(self._cache['ent'], dtype=float)

    def overlap(self, other):
        """<self|other> (self conjugated), MPS.overlap."""
        if self.L != other.L:
            raise ValueError('length mismatch')
        return self._ctx.overlap(0, other._ctx, 0) * self.norm * other.norm

    def correlation_function(self, op1, op2, sites1=None, sites2=None):
        s1 = list(range(self.L)) if sites1 is None else list(sites1)
        s2 = list(range(self.L)) if sites2 is None else list(sites2)
        out = np.zeros((len(s1), len(s2)), dtype=complex)
        for a, i in enumerate(s1):
            for b, j in enumerate(s2):
                out[a, b] = self._ctx.correlation(0, i, j, _as_matrix(op1, self.sites[i]),
                                                  _as_matrix(op2, self.sites[j]))
        return np.real_if_close(out)

    # ------------------------------------------------------------------ checkpoint / resume
    def save(self, path):
        """Write the state (site tensors in 'B' form, Schmidt values, site bookkeeping) to an ``.npz`` file.
        The reference keeps states only in Python lists (no checkpointing); this is the engine's
        resume format for long runs."""
        data = {'L': np.int64(self.L), 'norm': np.float64(self.norm),
                'conserve': np.array([str(getattr(s, 'conserve', None)) for s in self.sites]),
                'up': np.array([int(getattr(s, '_up', 0)) for s in self.sites], dtype=np.int64)}
        for i in range(self.L):
            data[f'B{i}'] = self.get_B(i, 'B')
        for b in range(self.L + 1):
            data[f'S{b}'] = self._ctx.get_S(0, b)
        np.savez(self._npz_path(path), **data)

    @staticmethod
    def _npz_path(path):
        """np.savez appends '.npz' to a name without it; save and load agree on the final name."""
        path = os.fspath(path)
        return path if path.endswith('.npz') else path + '.npz'

    @classmethod
    def load(cls, path, device=0):
        """Inverse of :meth:`save`."""
        with np.load(cls._npz_path(path), allow_pickle=False) as z:
            L = int(z['L'])
            Bs = [z[f'B{i}'] for i in range(L)]
            Ss = [z[f'S{b}'] for b in range(L + 1)]
            conserve, up, norm = [str(c) for c in z['conserve']], z['up'], float(z['norm'])
        cap = max(max(b.shape[0], b.shape[2]) for b in Bs)
        ctx = Context(L, cap, 1, device)
        for i, b in enumerate(Bs):
            ctx.set_site(0, i, b)
        for b, s_ in enumerate(Ss):
            ctx.set_S(0, b, s_)
        sites = []
        for c, u in zip(conserve, up):
            st = SpinHalfSite(None if c == 'None' else c)
            st._up = int(u)
            st.state_labels = {'up': st._up, 'down': 1 - st._up}
            sites.append(st)
        out = cls(ctx, sites)
        out.norm = norm
        return out

    # ------------------------------------------------------------------ misc
    def to_statevector(self):
        """Dense amplitudes (small L; host contraction of downloaded tensors, for tests/debugging)."""
        v = self.get_B(0, 'B')
        for i in range(1, self.L):
            v = np.tensordot(v, self.get_B(i, 'B'), axes=(v.ndim - 1, 0))
        return v.reshape(-1) * self.norm
