#!/usr/bin/env python3
"""
Generate tests/golden/reference_golden.json by running the UNMODIFIED reference
modules from /root/reference on top of the tenpy shim (oracle/tenpy_shim).

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no
/root/reference):

    python oracle/make_golden.py

What is pinned by these vectors: everything the reference's own Python does
(disorder fields, expm gates, gate order, kick, observable definitions, FFT
post-processing, DTC detector, config parser).  What is NOT pinned: TeNPy's
internals, which the shim restates (see oracle/__init__.py).
"""

import json
import os
import sys
import types
from unittest import mock

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('TC_REFERENCE_ROOT', '/root/reference')

sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, 'tenpy_shim'))
sys.path.insert(0, os.path.join(REF, 'src'))
sys.path.insert(0, REF)

# matplotlib is absent from the image and only used for plotting in main.py
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
    sys.modules.setdefault(name, mock.MagicMock(name=name))

from core.tensor_utils import create_initial_state, pauli_matrices, calculate_entanglement_entropy  # noqa: E402
from core import observables as robs  # noqa: E402
from models.kicked_ising import KickedIsingModel  # noqa: E402
from dynamics.tebd_evolution import CustomFloquet, TEBDEvolution  # noqa: E402
import main as rmain  # noqa: E402


def c2l(a):
    a = np.asarray(a)
    return {'re': a.real.tolist(), 'im': a.imag.tolist()}


def evolve_case(L, J, h, tau, seed, eps, n_periods, state='neel'):
    model = KickedIsingModel(L, J, h, tau, disorder_seed=seed)
    if eps != 0.0:
        # the only way to get an imperfect pulse through the reference API
        model.pi_pulse_gate = scipy.linalg.expm(-1j * np.pi / 2 * (1 - eps) * model.sigma_x)
    psi0 = create_initial_state(L, state)
    states, times, info = CustomFloquet(model).evolve_floquet(psi0, n_periods)
    Z = [[robs.magnetization(s, 'z', site=i) for i in range(L)] for s in states]
    return {
        'params': dict(L=L, J=J, h=h, tau=tau, seed=seed, eps=eps, n_periods=n_periods, state=state),
        'h_fields': model.h_fields.tolist(),
        'times': list(times),
        'Z': Z,
        'M_total': [robs.magnetization(s, 'z') for s in states],
        'M_stag': [robs.staggered_magnetization(s) for s in states],
        'LE': [robs.calculate_loschmidt_echo(psi0, s) for s in states],
        'S_ent': [s.entanglement_entropy().tolist() for s in states] if L > 1 else [],
        'chi': [list(s.chi) for s in states],
        'bond_dimensions': info['bond_dimensions'],
        'final_bond_dim': info['final_bond_dim'],
        'order_parameter': robs.order_parameter(states[-1], list(range(0, L, 2)), list(range(1, L, 2))),
        'SL_mid': np.asarray(robs.entanglement_spectrum(states[-1], L // 2)).tolist(),
        'Mx_final': robs.magnetization(states[-1], 'x'),
        'My_final': robs.magnetization(states[-1], 'y'),
        'corr_zz_0_3': c2l(robs.correlation_function(states[-1], 'z', 'z', 0, min(3, L - 1))),
        'corr_xy_1_2': c2l(robs.correlation_function(states[-1], 'x', 'y', min(1, L - 1), min(2, L - 1))),
        'participation_ratio': robs.participation_ratio(states[-1]),
    }


def main():
    out = {'generator': 'oracle/make_golden.py', 'reference': REF,
           'up_index': int(os.environ.get('TC_ORACLE_UP_INDEX', 1))}

    # ---- model construction (kicked_ising.py:35-98)
    models = []
    for (seed, h, L, J, tau) in [(42, 0.2, 4, 1.0, 1.0), (42, 0.25, 8, 1.0, 1.0), (123, 0.4, 4, 1.0, 0.5),
                                 (42, 0.3, 6, 0.7, 1.3)]:
        m = KickedIsingModel(L, J, h, tau, disorder_seed=seed)
        models.append({'seed': seed, 'h': h, 'L': L, 'J': J, 'tau': tau,
                       'h_fields': m.h_fields.tolist(),
                       'pi_pulse': c2l(m.pi_pulse_gate),
                       'ising_gates': [c2l(g) for g in m.ising_gates]})
    out['models'] = models
    m = KickedIsingModel(5, 1.0, 0.1, 1.0, bc='periodic', disorder_seed=7)
    out['periodic_gate_count'] = len(m.ising_gates)

    # ---- evolutions through the reference's own code path
    out['evolutions'] = [
        evolve_case(8, 1.0, 0.25, 1.0, 42, 0.0, 12),                  # chi == 1 laws
        evolve_case(4, 1.0, 0.2, 1.0, 42, 0.0, 6, state='all_up'),
        evolve_case(10, 1.0, 0.3, 1.0, 42, 0.1, 40),                   # SURVEY A.3 last bullet
        evolve_case(12, 1.0, 0.3, 1.0, 7, 0.05, 12),
        evolve_case(6, 0.8, 0.5, 0.7, 3, 0.2, 25, state='all_down'),
        evolve_case(1, 1.0, 0.3, 1.0, 5, 0.1, 3),                      # L = 1 edge case
        evolve_case(2, 1.0, 0.3, 1.0, 5, 0.1, 5),
        evolve_case(7, 1.0, 0.4, 0.9, 11, 0.15, 15),                   # odd L
    ]

    # ---- evolve()/floquet_step()/evolve_floquet_period bookkeeping
    model = KickedIsingModel(6, 1.0, 0.2, 1.0, disorder_seed=42)
    psi0 = create_initial_state(6, 'neel')
    states, times = model.evolve(psi0, 5)
    out['evolve_times'] = list(times)
    st2, t2, info = CustomFloquet(model).evolve_floquet(psi0, 7, measure_every=3)
    out['measure_every'] = {'n_states': len(st2), 'times': list(t2), 'keys': sorted(info.keys())}
    te = TEBDEvolution(model, dt=0.1, max_chi=50, trunc_params={'svd_min': 1e-10})
    out['tebd_trunc_params'] = te.trunc_params
    out['hamiltonian_terms_keys'] = sorted(model.get_hamiltonian_terms().keys())
    out['entropy_helper_cut2'] = float(calculate_entanglement_entropy(states[-1], 2))

    # ---- host FFT post-processing (observables.py:124-221, 254-277, 372-487)
    rng = np.random.default_rng(0)
    series = {}
    k = np.arange(31)
    series['alt31'] = ((-1.0) ** k).tolist()
    series['cos_half'] = (np.cos(np.pi * np.arange(64)) * np.exp(-0.01 * np.arange(64))
                          + 0.05 * rng.standard_normal(64)).tolist()
    series['cos_fund'] = np.cos(2 * np.pi * np.arange(50) * 2.0 / 2.0 / 2.0 * 0.5 + 0.3).tolist()
    series['noise'] = rng.standard_normal(81).tolist()
    series['le_decay'] = (0.5 + 0.5 * (-1.0) ** np.arange(81) * np.exp(-0.02 * np.arange(81))).tolist()
    series['short'] = [1.0, 0.0, 1.0, 0.0, 1.0]
    post = {}
    for name, s in series.items():
        s_arr = np.array(s)
        for period in (2.0, 4.0):
            t = np.arange(len(s)) * period
            key = f'{name}|T={period}'
            fund, sub = robs.subharmonic_response(list(s), period)
            ent = {
                'subharmonic_response': [float(fund), float(sub)],
                'extract_subharmonic_amplitude': robs.extract_subharmonic_amplitude(t, s_arr, period),
                'extract_from_loschmidt': robs.extract_subharmonic_amplitude_from_loschmidt(t, s_arr, period),
                'detect_period_doubling': float(robs.detect_period_doubling_from_loschmidt(list(np.abs(s_arr)))),
                'stringent_dtc_detection': float(rmain.stringent_dtc_detection(list(np.abs(s_arr)), list(t), period)),
            }
            if len(s) > 2:
                f, p = rmain.calculate_fourier_spectrum(t, s_arr, period)
                ent['fourier_freqs'] = np.asarray(f).tolist()
                ent['fourier_power'] = np.asarray(p).tolist()
                ent['fourier_peak_bin'] = int(np.argmax(p))
            post[key] = ent
    out['series'] = series
    out['post'] = post
    out['fidelity_decay'] = float(robs.fidelity_decay(list(np.exp(-0.05 * np.arange(20) * 2.0)),
                                                      list(np.arange(20) * 2.0)))

    # ---- config parser + phase point (main.py:39-130, 275-415)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        params = rmain.read_parameters('config.txt')
    finally:
        os.chdir(cwd)
    out['config_params'] = params
    pp = {}
    for (hj, tj) in [(0.2, 2.0), (0.7, 0.9), (0.0, 3.8)]:
        r = rmain.calculate_phase_point(hj, tj, params)
        pp[f'{hj}|{tj}'] = {k2: (bool(v) if isinstance(v, (bool, np.bool_)) else float(v)) for k2, v in r.items()}
    out['phase_points'] = pp

    # ---- initial states (tensor_utils.py:28-62)
    init = {}
    for st in ('all_up', 'all_down', 'neel'):
        p = create_initial_state(4, st)
        init[st] = {'Z': [robs.magnetization(p, 'z', site=i) for i in range(4)], 'chi': list(p.chi),
                    'norm': float(p.norm)}
    np.random.seed(99)
    p = create_initial_state(6, 'random')
    init['random_seed99'] = {'Z': [robs.magnetization(p, 'z', site=i) for i in range(6)]}
    out['initial_states'] = init
    out['pauli'] = {k2: c2l(v) for k2, v in pauli_matrices().items()}

    dst = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(dst, exist_ok=True)
    path = os.path.join(dst, 'reference_golden.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
