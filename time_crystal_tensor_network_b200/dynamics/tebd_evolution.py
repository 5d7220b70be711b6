"""Drop-in for the reference's ``src/dynamics/tebd_evolution.py`` on the B200 engine.

``CustomFloquet`` is the live driver (tebd_evolution.py:199-259): a loop over
``model.floquet_step`` with snapshots and bookkeeping.  ``TEBDEvolution`` keeps the reference's
constructor and method names; ``evolve_floquet_period`` is the live alias (tebd_evolution.py:178-188).
Its ``evolve`` (dead code in the reference: it calls a TEBDEngine method that does not exist) is
implemented here as what its docstring promises: second-order Trotter real-time TEBD under the
model's nearest-neighbour Hamiltonian with TeNPy ``truncate()`` semantics.
"""
import time
from typing import Dict, List, Tuple

import numpy as np
import scipy.linalg

from ..engine import EngineError
from ..mps import MPS, CHI_HARD_CAP


class CustomFloquet:
    """Floquet evolution of a KickedIsingModel with per-period snapshots."""

    def __init__(self, kicked_ising_model, trunc_params: Dict = None):
        self.model = kicked_ising_model
        self.trunc_params = trunc_params if trunc_params is not None else \
            {'chi_max': 100, 'svd_min': 1e-12, 'trunc_cut': 1e-10}

    def evolve_floquet(self, psi_initial: MPS, n_periods: int,
                       measure_every: int = 1) -> Tuple[List[MPS], List[float], Dict]:
        """Returns (states, times, info); a state is stored after period p when p % measure_every == 0
        (0-based), times are (p + 1) * 2 tau (tebd_evolution.py:218-259)."""
        top = lambda psi: max(psi.chi) if psi.chi else 1
        # one working state advances in place when the model offers it (KickedIsingModel.floquet_step_inplace: the same
        # arithmetic as floquet_step without a new context per period); the returned states are snapshots without SVD
        # workspace.  Any other model object goes through its floquet_step as in the reference.
        snap = lambda psi: psi.copy(storage=True) if isinstance(psi, MPS) else psi.copy()
        step = getattr(self.model, 'floquet_step_inplace', None) or self.model.floquet_step
        states, times, bond_dims = [snap(psi_initial)], [0.0], [top(psi_initial)]
        psi = psi_initial.copy()
        t0 = time.time()
        for period in range(n_periods):
            psi = step(psi, self.trunc_params)
            if period % measure_every == 0:
                states.append(snap(psi))
                times.append((period + 1) * 2 * self.model.tau)
                bond_dims.append(top(psi))
        wall = time.time() - t0
        info = {
            'wall_time': wall,
            'bond_dimensions': bond_dims,
            'periods_per_second': n_periods / wall if wall > 0 else float('inf'),
            'final_bond_dim': top(psi),
            'n_periods': n_periods,
        }
        return states, times, info


class TEBDEvolution:
    """Real-time TEBD wrapper with the reference's interface (tebd_evolution.py:18-188)."""

    def __init__(self, model, dt: float = 0.1, max_chi: int = 100, trunc_params: Dict = None):
        self.model, self.dt, self.max_chi = model, dt, max_chi
        if trunc_params is None:
            self.trunc_params = {'chi_max': max_chi, 'svd_min': 1e-12, 'trunc_cut': 1e-10}
        else:
            self.trunc_params = trunc_params
            self.trunc_params.setdefault('chi_max', max_chi)

    # ------------------------------------------------------------------ Hamiltonian -> bond gates
    def _bond_terms(self):
        """Hermitian 4x4 bond terms H_b with sum_b H_b = H.  Accepts a model exposing ``H_bond`` (list of
        4x4 arrays) or a KickedIsingModel (static Ising part J zz + h_i z, each field split evenly
        between the bonds that contain the site)."""
        m = self.model
        if hasattr(m, 'H_bond'):
            return [np.asarray(h, dtype=complex).reshape(4, 4) for h in m.H_bond]
        if hasattr(m, 'h_fields') and hasattr(m, 'J'):
            L = m.n_sites
            z, one = np.diag([1.0, -1.0]).astype(complex), np.eye(2, dtype=complex)
            terms = []
            for i in range(L - 1):
                wl = 1.0 if i == 0 else 0.5
                wr = 1.0 if i + 1 == L - 1 else 0.5
                terms.append(m.J * np.kron(z, z) + wl * m.h_fields[i] * np.kron(z, one)
                             + wr * m.h_fields[i + 1] * np.kron(one, z))
            return terms
        raise TypeError('model must provide H_bond (list of 4x4 bond Hamiltonians) or be a KickedIsingModel')

    def suzuki_trotter_gates(self, hamiltonian_terms: Dict, dt: float) -> List[np.ndarray]:
        """exp(-i dt H_term) for every entry except 'single_site_terms' (tebd_evolution.py:128-149)."""
        return [scipy.linalg.expm(-1j * dt * op) for name, op in hamiltonian_terms.items()
                if name != 'single_site_terms']

    def _trotter(self, psi_initial: MPS, n_steps: int, prefactor, observe_every: int, dt: float):
        """n_steps second-order Trotter steps G_even(dt/2) G_odd(dt) G_even(dt/2) with bond gates
        expm(prefactor * dt * H_b): prefactor = -1j is real time, -1 is imaginary time (the state is
        renormalised by every update; the O(dt) loss of canonical form under non-unitary gates is repaired by the
        caller, imaginary_time_evolution, with MPS.canonical_form)."""
        terms = self._bond_terms()
        L = psi_initial.L
        if len(terms) != L - 1:
            raise ValueError('need one bond term per nearest-neighbour bond')
        gates = np.array([scipy.linalg.expm(prefactor * dt * (0.5 if b % 2 == 0 else 1.0) * h)
                          for b, h in enumerate(terms)]).reshape(1, L - 1, 4, 4) if L > 1 else None
        tp = self.trunc_params
        chi_max = int(tp.get('chi_max') or 0)
        cap = min(2 ** (L // 2), chi_max if chi_max > 0 else CHI_HARD_CAP, CHI_HARD_CAP)
        psi = psi_initial.copy(chi_cap=max(cap, 1))
        psi._ctx.set_model(gates, np.eye(2, dtype=complex).reshape(1, 2, 2))
        psi._ctx.set_trunc('tebd', chi_max=chi_max, svd_min=tp.get('svd_min') or 0.0,
                           trunc_cut=tp.get('trunc_cut') or 0.0)
        psi._ctx.trunc_err(reset=True)
        states, times = [psi_initial.copy(storage=True)], [0.0]
        bond_dims, entropies, errs = [psi_initial.chi], [psi_initial.entanglement_entropy()], []
        t0 = time.time()
        for step in range(n_steps):
            if L > 1:
                psi._ctx.apply_layer(0, 0)
                if L > 2:
                    psi._ctx.apply_layer(1, 0)
                psi._ctx.apply_layer(0, 0)
            psi._touch()
            if step % observe_every == 0:
                states.append(psi.copy(storage=True))
                times.append((step + 1) * dt)
                bond_dims.append(psi.chi)
                entropies.append(psi.entanglement_entropy())
                errs.append(float(psi._ctx.trunc_err()[0]))
        wall = time.time() - t0
        fl = psi._ctx.flags()
        if fl['chi_cap_overflow'] or fl['svd_not_converged']:
            raise EngineError(f'TEBD failed on the device: {fl}')
        info = {
            'wall_time': wall,
            'bond_dimensions': bond_dims,
            'entanglement_entropies': entropies,
            'truncation_errors': errs,
            'final_bond_dim': psi.chi,
            'n_steps': n_steps,
        }
        return states, times, info, psi

    def evolve(self, psi_initial: MPS, total_time: float, observe_every: int = 1) -> Tuple[List[MPS], List[float], Dict]:
        """Real-time second-order Trotter TEBD under the model's nearest-neighbour Hamiltonian; returns
        (states, times, info) with the reference's info keys (tebd_evolution.py:51-108)."""
        states, times, info, _ = self._trotter(psi_initial, int(total_time / self.dt), -1j, observe_every, self.dt)
        return states, times, info

    def imaginary_time_evolution(self, psi_initial: MPS, dts=(0.1, 0.05, 0.02, 0.01), steps_per_dt: int = 100,
                                 recanonicalize: bool = True) -> Tuple[MPS, Dict]:
        """Ground-state preparation by imaginary-time TEBD (the README's claim, BASELINE config 5): for every dt of
        the schedule, ``steps_per_dt`` second-order steps of exp(-dt H).  Non-unitary gates leave the chain out of
        canonical form at O(dt); every stage therefore ends with ``MPS.canonical_form()`` (identity-gate sweeps, the
        state itself unchanged), so that the energies, the Schmidt values and whatever evolution follows start from an
        exactly canonical chain.  ``recanonicalize=False`` keeps the plain TEBD behaviour (what TeNPy's TEBDEngine
        does by itself).  Returns (psi, info) with the energy after every stage (sum of <H_b>)."""
        psi = psi_initial
        energies, chis, sweeps = [], [], []
        for dt in dts:
            _, _, _, psi = self._trotter(psi, steps_per_dt, -1.0, max(steps_per_dt, 1), dt)
            if recanonicalize:
                sweeps.append(psi.canonical_form())
            energies.append(self.energy(psi))
            chis.append(max(psi.chi) if psi.chi else 1)
        return psi, {'energies': energies, 'bond_dimensions': chis, 'dts': list(dts), 'canonical_sweeps': sweeps}

    def energy(self, psi: MPS) -> float:
        """<H> = sum_b <H_b> through the Pauli expansion of every bond term (two-point correlators on the device)."""
        paulis = [np.eye(2, dtype=complex), np.array([[0, 1], [1, 0]], dtype=complex),
                  np.array([[0, -1j], [1j, 0]], dtype=complex), np.array([[1, 0], [0, -1]], dtype=complex)]
        total = 0.0
        for b, h in enumerate(self._bond_terms()):
            h4 = h.reshape(2, 2, 2, 2)
            for ia, pa in enumerate(paulis):
                for ib, pb in enumerate(paulis):
                    c = np.einsum('pqrs,rp,sq->', h4, pa, pb) / 4.0        # tr[(pa x pb) H_b] / 4
                    if abs(c) < 1e-15:
                        continue
                    if ia == 0 and ib == 0:
                        total += c.real
                    elif ia == 0:
                        total += (c * psi.expectation_value(pb, sites=[b + 1])[0]).real
                    elif ib == 0:
                        total += (c * psi.expectation_value(pa, sites=[b])[0]).real
                    else:
                        total += (c * psi.correlation_function(pa, pb, sites1=[b], sites2=[b + 1])[0, 0]).real
        return float(total)

    def real_time_evolution(self, psi_initial: MPS, hamiltonian, total_time: float,
                            observe_every: int = 1) -> Tuple[List[MPS], List[float], Dict]:
        """As in the reference, the ``hamiltonian`` argument is not consulted (tebd_evolution.py:110-126)."""
        return self.evolve(psi_initial, total_time, observe_every)

    def benchmark_performance(self, psi_initial: MPS, n_steps: int = 100) -> Dict:
        """Wall-clock benchmark with the reference's result keys (tebd_evolution.py:151-176)."""
        t0 = time.time()
        _, _, info = self.evolve(psi_initial, n_steps * self.dt, observe_every=n_steps)
        wall = time.time() - t0
        errs = info['truncation_errors']
        return {
            'wall_time': wall,
            'steps_per_second': n_steps / wall,
            'final_bond_dim': info['final_bond_dim'],
            'memory_usage': sum(sum(b) if hasattr(b, '__len__') else b for b in info['bond_dimensions']) * 8 / 1024 ** 2,
            'truncation_error': errs[-1] if errs else 0,
        }

    def evolve_floquet_period(self, psi: MPS) -> MPS:
        """One Floquet period through the model (tebd_evolution.py:178-188)."""
        return self.model.floquet_step(psi, self.trunc_params)
