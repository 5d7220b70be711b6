"""Drop-in for the reference's ``src/core`` package (same public names)."""
from .tensor_utils import create_initial_state, pauli_matrices, apply_two_site_gate
from .observables import calculate_loschmidt_echo, magnetization, correlation_function

__all__ = ['create_initial_state', 'pauli_matrices', 'apply_two_site_gate',
           'calculate_loschmidt_echo', 'magnetization', 'correlation_function']
