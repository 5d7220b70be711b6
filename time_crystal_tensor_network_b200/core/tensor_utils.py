"""Drop-in for the reference's ``src/core/tensor_utils.py``: product-state factory, Pauli matrices
and thin MPS helpers, on the GPU engine instead of TeNPy."""
import numpy as np

from ..mps import MPS, SpinHalfSite

_PAULI = {
    'I': [[1, 0], [0, 1]],
    'X': [[0, 1], [1, 0]],
    'Y': [[0, -1j], [1j, 0]],
    'Z': [[1, 0], [0, -1]],
}


def pauli_matrices():
    """{'I','X','Y','Z'} -> complex 2x2 arrays (tensor_utils.py:13-25)."""
    return {k: np.array(v, dtype=complex) for k, v in _PAULI.items()}


def create_initial_state(n_sites, state_type="all_up"):
    """Z-basis product MPS (tensor_utils.py:28-62).  'random' draws from the global NumPy RNG, one
    ``np.random.choice`` per site, so seeding behaves as in the reference."""
    sites = [SpinHalfSite(conserve='parity') for _ in range(n_sites)]
    if state_type == "all_up":
        labels = ["up"] * n_sites
    elif state_type == "all_down":
        labels = ["down"] * n_sites
    elif state_type == "neel":
        labels = ["down" if i % 2 else "up" for i in range(n_sites)]
    elif state_type == "random":
        labels = [np.random.choice(["up", "down"]) for _ in range(n_sites)]
    else:
        raise ValueError(f"Unknown state type: {state_type}")
    return MPS.from_product_state(sites, labels, bc='finite')


def apply_two_site_gate(psi, gate, i, j, trunc_params=None):
    """4x4 gate on adjacent sites (tensor_utils.py:65-105); returns a new MPS, input untouched.
    As in the reference, ``trunc_params`` is accepted and not used: the update keeps every
    singular value above 1e-13."""
    if abs(i - j) != 1:
        raise ValueError("Sites must be adjacent for two-site gate")
    out = psi.copy()
    out.apply_local_op(min(i, j), np.asarray(gate).reshape(2, 2, 2, 2), unitary=True)
    return out


def create_time_evolution_gates(J, h, tau, n_sites):
    """Reproduces tensor_utils.py:108-142 literally, including its element-wise ``np.exp`` (this is
    not a matrix exponential; nothing calls it)."""
    p = pauli_matrices()
    h2 = J * np.kron(p['Z'], p['Z']) + h * np.kron(p['Z'], p['I']) + h * np.kron(p['I'], p['Z'])
    return {'ising_evolution': np.exp(-1j * tau * h2), 'pi_pulse': np.exp(-1j * np.pi / 2 * p['X'])}


def measure_magnetization(psi, direction='z'):
    """Total magnetisation, summed site by site (tensor_utils.py:145-166)."""
    op = pauli_matrices()[direction.upper()]
    total = 0.0
    for i in range(psi.L):
        total += psi.expectation_value(op, sites=[i]).real
    return total


def calculate_entanglement_entropy(psi, cut):
    """Entropy of the bond to the right of site ``cut`` (tensor_utils.py:169-180)."""
    return psi.entanglement_entropy()[cut]


def mps_overlap(psi1, psi2):
    """<psi1|psi2> (tensor_utils.py:183-192)."""
    return psi1.overlap(psi2)
