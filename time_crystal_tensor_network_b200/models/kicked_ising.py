"""Drop-in for the reference's ``src/models/kicked_ising.py`` on the B200 engine.

U_F = U_Ising(tau/2) . prod_j exp(-i pi/2 sigma^x_j) . U_Ising(tau/2), with
U_Ising = odd-bond gates . even-bond gates, every two-site gate followed by its own SVD
(kicked_ising.py:100-160).  The whole period runs on the GPU as four batched layer launches; the
host only builds the 4x4 / 2x2 matrices (kicked_ising.py:73-98) and owns the handles.

Truncation: the reference passes ``trunc_params`` around but its gate routine never reads them
(kicked_ising.py:162-188); the effective rule is TeNPy's ``apply_local_op`` default (keep
sigma > 1e-13, renormalise).  ``truncation = 'reference'`` (default) reproduces that.  Setting
``model.truncation = 'tebd'`` makes ``chi_max`` / ``svd_min`` / ``trunc_cut`` effective with TeNPy's
``truncate()`` semantics.
"""
from typing import Dict, List, Tuple

import numpy as np
import scipy.linalg

from ..engine import EngineError
from ..mps import MPS, SpinHalfSite, CHI_HARD_CAP


class KickedIsingModel:
    """Floquet kicked-Ising chain with random longitudinal fields."""

    def __init__(self, n_sites: int, J: float, h_disorder: float, tau: float,
                 bc: str = 'open', disorder_seed: int = None):
        self.n_sites, self.J, self.h_disorder, self.tau, self.bc = n_sites, J, h_disorder, tau, bc
        # the reference reseeds the *global* legacy RNG (kicked_ising.py:55-59); later draws
        # (e.g. create_initial_state(..., "random")) continue from this stream
        if disorder_seed is not None:
            np.random.seed(disorder_seed)
        self.h_fields = np.random.uniform(-h_disorder, h_disorder, n_sites)
        self.sites = [SpinHalfSite(conserve=None) for _ in range(n_sites)]
        self.sigma_x = np.array([[0, 1], [1, 0]], dtype=complex)
        self.sigma_y = np.array([[0, -1j], [1j, 0]], dtype=complex)
        self.sigma_z = np.array([[1, 0], [0, -1]], dtype=complex)
        self.sigma_I = np.eye(2, dtype=complex)
        self.truncation = 'reference'
        self._prepare_gates()

    # ------------------------------------------------------------------ gates (host, tiny)
    def _bond_hamiltonian(self, h_left, h_right):
        zz = np.kron(self.sigma_z, self.sigma_z)
        return (self.J * zz + h_left * np.kron(self.sigma_z, self.sigma_I)
                + h_right * np.kron(self.sigma_I, self.sigma_z))

    def _prepare_gates(self):
        """pi-pulse and per-bond Ising gates; each bulk site's field enters both neighbouring bond
        gates, exactly as in the reference (kicked_ising.py:83-85)."""
        self.pi_pulse_gate = scipy.linalg.expm(-1j * np.pi / 2 * self.sigma_x)
        half = -1j * self.tau / 2
        self.ising_gates = [scipy.linalg.expm(half * self._bond_hamiltonian(self.h_fields[i], self.h_fields[i + 1]))
                            for i in range(self.n_sites - 1)]
        if self.bc == 'periodic' and self.n_sites > 2:
            self.ising_gates.append(
                scipy.linalg.expm(half * self._bond_hamiltonian(self.h_fields[-1], self.h_fields[0])))

    # ------------------------------------------------------------------ device plumbing
    def _trunc_settings(self, trunc_params):
        if self.truncation == 'tebd':
            tp = trunc_params or {}
            return dict(mode='tebd', chi_max=int(tp.get('chi_max') or 0), svd_min=tp.get('svd_min') or 0.0,
                        trunc_cut=tp.get('trunc_cut') or 0.0)
        return dict(mode='reference', cutoff=1e-13)

    def _room_for_one_period(self, psi, settings):
        """Bond dimension the context must be able to hold after one period: each bond is updated twice and an
        update multiplies it by at most the operator-Schmidt rank of its gate (2 for the diagonal Ising gates
        _prepare_gates builds, up to 4 for general 4x4 matrices a user assigned to ``ising_gates``)."""
        L = psi.L
        now = max(psi._chi_full())
        diag = all(np.count_nonzero(g - np.diag(np.diag(g))) == 0 for g in self.ising_gates)
        cap = min(2 ** (L // 2), (4 if diag else 16) * now)
        if settings['mode'] == 'tebd' and settings['chi_max'] > 0:
            cap = min(cap, settings['chi_max'])
        cap = max(cap, now, 1)
        if cap > CHI_HARD_CAP:
            raise EngineError(f'bond dimension would exceed TC_CHI_HARD_CAP={CHI_HARD_CAP}; '
                              "use model.truncation = 'tebd' with a chi_max")
        return cap

    def _load(self, psi, settings):
        if psi.L != self.n_sites:
            raise ValueError('state and model have different lengths')
        if len(self.ising_gates) > max(self.n_sites - 1, 0):
            # the reference hands the wrap-around gate to apply_local_op(L-1, two-site op), which
            # cannot fit on a finite MPS (kicked_ising.py:136,186)
            raise ValueError('local operator does not fit on finite MPS')
        gates = np.array(self.ising_gates, dtype=complex).reshape(1, -1, 4, 4) if self.n_sites > 1 else None
        psi._ctx.set_model(gates, np.asarray(self.pi_pulse_gate, dtype=complex).reshape(1, 2, 2))
        psi._ctx.set_trunc(**settings)

    @staticmethod
    def _check_flags(psi):
        fl = psi._ctx.flags()
        if fl['chi_cap_overflow'] or fl['svd_not_converged']:
            raise EngineError(f'update failed on the device: {fl}')

    def _advance(self, psi, settings):
        """One period in place on ``psi`` (which this module owns)."""
        psi._grow(self._room_for_one_period(psi, settings))
        self._load(psi, settings)
        psi._ctx.floquet_step(1)
        psi._touch()
        self._check_flags(psi)

    def floquet_step_inplace(self, psi: MPS, trunc_params: Dict = None) -> MPS:
        """One Floquet period applied to ``psi`` itself (no new context): what the time loops of ``evolve`` and
        ``CustomFloquet.evolve_floquet`` use on their private working copy.  Same arithmetic as ``floquet_step``."""
        if trunc_params is None:
            trunc_params = {'chi_max': 100, 'svd_min': 1e-12}
        self._advance(psi, self._trunc_settings(trunc_params))
        return psi

    # ------------------------------------------------------------------ public API
    def floquet_step(self, psi: MPS, trunc_params: Dict = None) -> MPS:
        """One Floquet period; returns a new MPS and leaves ``psi`` untouched (kicked_ising.py:100-126)."""
        if trunc_params is None:
            trunc_params = {'chi_max': 100, 'svd_min': 1e-12}
        settings = self._trunc_settings(trunc_params)
        out = psi.copy(chi_cap=self._room_for_one_period(psi, settings))
        self._advance(out, settings)
        return out

    def _apply_ising_evolution(self, psi: MPS, trunc_params: Dict) -> MPS:
        """Even bonds then odd bonds (kicked_ising.py:128-148); new MPS."""
        settings = self._trunc_settings(trunc_params)
        out = psi.copy(chi_cap=self._room_for_one_period(psi, settings))
        self._load(out, settings)
        out._ctx.apply_layer(0, 0)
        if self.n_sites > 2:
            out._ctx.apply_layer(1, 0)
        out._touch()
        self._check_flags(out)
        return out

    def _apply_pi_pulse(self, psi: MPS, trunc_params: Dict = None) -> MPS:
        """Kick on every site (kicked_ising.py:150-160); new MPS."""
        out = psi.copy()
        out._ctx.set_model(None, np.asarray(self.pi_pulse_gate, dtype=complex).reshape(1, 2, 2))
        out._ctx.apply_kick()
        out._touch()
        return out

    def _apply_two_site_gate(self, psi: MPS, gate: np.ndarray, bond_idx, trunc_params: Dict = None) -> MPS:
        i = bond_idx if isinstance(bond_idx, (int, np.integer)) else bond_idx[0]
        out = psi.copy()
        out.apply_local_op(i, np.asarray(gate).reshape(2, 2, 2, 2), unitary=True)
        return out

    def _apply_single_site_gate(self, psi: MPS, gate: np.ndarray, site: int) -> MPS:
        out = psi.copy()
        out.apply_local_op(site, gate, unitary=True)
        return out

    def evolve(self, psi_initial: MPS, n_steps: int, trunc_params: Dict = None) -> Tuple[List[MPS], List[float]]:
        """n_steps periods; returns ([psi(0), ..., psi(n)], [0, 2 tau, ...]) (kicked_ising.py:210-239)."""
        if trunc_params is None:
            trunc_params = {'chi_max': 100, 'svd_min': 1e-12}
        settings = self._trunc_settings(trunc_params)
        # one working context advances in place; the returned states are snapshots without SVD workspace
        states, times = [psi_initial.copy(storage=True)], [0.0]
        work = psi_initial.copy()
        for step in range(n_steps):
            self._advance(work, settings)
            states.append(work.copy(storage=True))
            times.append((step + 1) * 2 * self.tau)
        return states, times

    def get_hamiltonian_terms(self) -> Dict[str, np.ndarray]:
        return {'J': self.J, 'h_fields': self.h_fields, 'tau': self.tau,
                'pi_pulse': self.pi_pulse_gate, 'ising_gates': self.ising_gates}

    def calculate_phase_diagram_point(self, psi_initial: MPS, n_steps: int = 200,
                                      trunc_params: Dict = None) -> Dict[str, float]:
        """Observables of one phase-diagram point (kicked_ising.py:256-303)."""
        from ..core.observables import (calculate_loschmidt_echo, magnetization, subharmonic_response,
                                        order_parameter)
        states, _ = self.evolve(psi_initial, n_steps, trunc_params)
        echoes = [calculate_loschmidt_echo(psi_initial, s) for s in states]
        mags = [magnetization(s, 'z') for s in states]
        fund, sub = subharmonic_response(mags, 2 * self.tau)
        last = states[-1]
        return {
            'loschmidt_echo_final': echoes[-1],
            'subharmonic_amplitude': sub,
            'fundamental_amplitude': fund,
            'order_parameter': order_parameter(last, list(range(0, self.n_sites, 2)),
                                               list(range(1, self.n_sites, 2))),
            'max_bond_dimension': max(last.chi) if last.chi else 1,
            'final_magnetization': mags[-1],
        }
