"""Latency of the reference-shaped single-chain call: CustomFloquet.evolve_floquet at L = 16 (chi_max = 64) and L = 20
(chi_max = 128), eps = 0.1, TEBD truncation.  One chain gives a layer launch 8-10 matrices: the question is whether
the SVD kernels should then spread a matrix over a thread-block cluster (TC_JACOBI=team) instead of one CTA per matrix.
usage: [TC_JACOBI=team] python scripts/single_chain_latency.py"""
import os
import sys
import time

import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from time_crystal_tensor_network_b200.models.kicked_ising import KickedIsingModel  # noqa: E402
from time_crystal_tensor_network_b200.dynamics.tebd_evolution import CustomFloquet  # noqa: E402
from time_crystal_tensor_network_b200.core.tensor_utils import create_initial_state  # noqa: E402
from time_crystal_tensor_network_b200.core.observables import magnetization  # noqa: E402

for L, chi, n in ((16, 64, 30), (20, 128, 30), (16, 64, 30), (20, 128, 30)):
    model = KickedIsingModel(L, 1.0, 0.3, 1.0, disorder_seed=42)
    model.pi_pulse_gate = scipy.linalg.expm(-1j * np.pi / 2 * 0.9 * model.sigma_x)
    model.truncation = 'tebd'
    psi0 = create_initial_state(L, 'neel')
    t0 = time.perf_counter()
    states, times, info = CustomFloquet(model, dict(chi_max=chi, svd_min=1e-12, trunc_cut=1e-7)).evolve_floquet(psi0, n)
    z = magnetization(states[-1], 'z')
    dt = time.perf_counter() - t0
    print({'L': L, 'chi_max': chi, 'periods': n, 'seconds': round(dt, 3), 'ms_per_period': round(1e3 * dt / n, 2),
           'final_bond_dim': int(info['final_bond_dim']), 'Z_mean': float(np.mean(z)),
           'variant': {k: v for k, v in os.environ.items() if k.startswith('TC_')}}, flush=True)
