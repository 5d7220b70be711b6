"""The assertions of the reference's own suites (tests/test_basic_functionality.py,
test_physics_validation.py, test_performance.py), restated for pytest and run against the drop-in modules on
the GPU.  Import paths are the reference's: only <repo>/src is put on sys.path."""
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'src'))
sys.path.insert(0, ROOT)

from core.tensor_utils import create_initial_state, pauli_matrices  # noqa: E402
from core.observables import (calculate_loschmidt_echo, magnetization, staggered_magnetization,  # noqa: E402
                              extract_subharmonic_amplitude, order_parameter)
from models.kicked_ising import KickedIsingModel  # noqa: E402
from dynamics.tebd_evolution import CustomFloquet, TEBDEvolution  # noqa: E402


def test_pauli_and_states():
    p = pauli_matrices()
    for a in 'XYZ':
        assert np.allclose(p[a] @ p[a], p['I'])
    assert np.allclose(p['X'] @ p['Y'] - p['Y'] @ p['X'], 2j * p['Z'])
    for st in ('all_up', 'all_down', 'neel', 'random'):
        psi = create_initial_state(6, st)
        assert psi.L == 6 and abs(psi.norm - 1.0) < 1e-10
    assert create_initial_state(1, 'all_up').L == 1
    with pytest.raises(ValueError):
        create_initial_state(4, 'invalid')


def test_model_construction():
    m = KickedIsingModel(n_sites=4, J=1.0, h_disorder=0.2, tau=1.0, disorder_seed=42)
    assert m.n_sites == 4 and m.J == 1.0 and m.tau == 1.0
    assert len(m.h_fields) == 4 and np.all(np.abs(m.h_fields) <= 0.2)
    assert len(m.ising_gates) == 3 and m.pi_pulse_gate.shape == (2, 2)
    m2 = KickedIsingModel(n_sites=4, J=1.0, h_disorder=0.2, tau=1.0, disorder_seed=43)
    assert not np.allclose(m.h_fields, m2.h_fields)
    t0 = time.time()
    for s in range(10):
        KickedIsingModel(8, 1.0, 0.3, 1.0, disorder_seed=s)
    assert time.time() - t0 < 5.0
    for h, tau in ((1e-10, 1.0), (10.0, 1.0), (0.2, 1e-3)):
        mm = KickedIsingModel(4, 1.0, h, tau, disorder_seed=1)
        assert abs(mm.floquet_step(create_initial_state(4, 'neel')).norm - 1.0) < 1e-8


def test_observables_on_product_states():
    up, dn, neel = (create_initial_state(4, s) for s in ('all_up', 'all_down', 'neel'))
    assert abs(calculate_loschmidt_echo(up, up) - 1.0) < 1e-10
    assert abs(calculate_loschmidt_echo(up, dn)) < 1e-10
    mu, md = magnetization(up, 'z'), magnetization(dn, 'z')
    assert abs(abs(mu) - 4.0) < 1e-8 and abs(mu + md) < 1e-8
    assert abs(magnetization(neel, 'z')) < 1e-8
    assert abs(abs(magnetization(up, 'z', site=0)) - 1.0) < 1e-8
    assert abs(staggered_magnetization(neel)) > 0.5 and abs(staggered_magnetization(up)) < 1e-8
    assert isinstance(magnetization(up, 'x'), float) and isinstance(calculate_loschmidt_echo(up, neel), float)
    assert up.expectation_value('Sz', sites=[0])[0].real == pytest.approx(0.5)


def test_physical_bounds_after_an_even_number_of_perfect_kicks():
    """The reference's physical-bounds case (tests/test_physics_validation.py:195-220): L = 6, tau = 0.8, weak disorder,
    ten periods with a perfect pi pulse -- the product state is back on the Neel state, the echo is 1 in exact arithmetic
    and must not come out above 1.0; magnetisations stay inside their bounds."""
    model = KickedIsingModel(n_sites=6, J=1.0, h_disorder=0.2, tau=0.8, disorder_seed=42)
    psi = create_initial_state(6, 'neel')
    for _ in range(10):
        psi = model.floquet_step(psi)
    for d in ('x', 'y', 'z'):
        assert abs(magnetization(psi, d)) <= 6.0 + 1e-9
        for site in range(3):
            assert abs(magnetization(psi, d, site=site)) <= 1.0 + 1e-9
    le = calculate_loschmidt_echo(create_initial_state(6, 'neel'), psi)
    assert 0.0 <= le <= 1.0 and le > 1.0 - 1e-12


def test_evolution_shapes_norms_and_timing():
    for L, limit in ((8, 0.1), (12, 0.5), (16, 2.0)):
        m = KickedIsingModel(L, 1.0, 0.3, 1.0, disorder_seed=42)
        psi = create_initial_state(L, 'neel')
        m.floquet_step(psi)                                   # first call pays context creation
        t0 = time.time()
        out = m.floquet_step(psi)
        assert time.time() - t0 < limit
        assert out.L == L and abs(out.norm - 1.0) < 1e-10
    m = KickedIsingModel(6, 1.0, 0.3, 0.5, disorder_seed=42)
    psi = create_initial_state(6, 'neel')
    cur = psi
    for _ in range(10):
        cur = m.floquet_step(cur)
        assert abs(cur.norm - psi.norm) < 1e-8
    states, times = m.evolve(psi, 10)
    assert len(states) == 11 and np.allclose(times, [i * 2 * m.tau for i in range(11)])
    st, tm, info = CustomFloquet(m, {'chi_max': 16, 'svd_min': 1e-12, 'trunc_cut': 1e-8}).evolve_floquet(psi, 5)
    assert len(st) == 6 and {'wall_time', 'bond_dimensions', 'final_bond_dim'} <= set(info)
    assert np.allclose(tm, [i * 2 * m.tau for i in range(6)])
    t5 = time.time(); m.evolve(psi, 5); t5 = time.time() - t5
    t20 = time.time(); m.evolve(psi, 20); t20 = time.time() - t20
    assert t20 / max(t5, 1e-3) < 8


def test_time_crystal_signatures():
    m = KickedIsingModel(8, 1.0, 0.25, 1.0, disorder_seed=42)
    psi0 = create_initial_state(8, 'neel')
    states, times = m.evolve(psi0, 20)
    stag = np.array([staggered_magnetization(s) for s in states])
    le = [calculate_loschmidt_echo(psi0, s) for s in states]
    assert np.std(stag) > 0.01 and le[-1] > 0.0
    assert all(-1e-10 <= x <= 1 + 1e-10 for x in le)
    assert extract_subharmonic_amplitude(np.array(times), stag, 2 * m.tau) > 0.1
    assert order_parameter(states[-1], [0, 2, 4, 6], [1, 3, 5, 7]) > 0.5
    dims = []
    for L in (4, 8, 12):
        mm = KickedIsingModel(L, 1.0, 0.3, 1.0, disorder_seed=42)
        _, _, info = CustomFloquet(mm).evolve_floquet(create_initial_state(L, 'neel'), 5)
        dims.append(info['final_bond_dim'])
    assert dims == sorted(dims)


def test_scaling_ceilings():
    for L in (16, 20, 24):
        m = KickedIsingModel(L, 1.0, 0.3, 1.0, disorder_seed=42)
        t0 = time.time()
        CustomFloquet(m, {'chi_max': 64, 'svd_min': 1e-12}).evolve_floquet(create_initial_state(L, 'neel'), 5)
        assert time.time() - t0 < 30
    m = KickedIsingModel(12, 1.0, 0.3, 1.0, disorder_seed=42)
    t0 = time.time()
    m.evolve(create_initial_state(12, 'neel'), 50)
    assert time.time() - t0 < 60
    te = TEBDEvolution(m)
    assert te.evolve_floquet_period(create_initial_state(12, 'neel')).L == 12


def test_perfect_dtc_driver_full_size():
    """basic:454-479 runs simulate_perfect_dtc at its real size (L = 64, 200 periods)."""
    import main as mm
    params = {'J': 1.0, 'CHI_MAX': 256, 'SVD_MIN': 1e-12, 'SVD_CUTOFF': 1e-7, 'RANDOM_SEED': 42}
    t0 = time.time()
    times, stag, total = mm.simulate_perfect_dtc(params)
    assert len(times) == len(stag) == len(total) == 201
    assert all(abs(s) <= 1 + 1e-10 for s in stag) and np.std(stag) > 0.01
    assert time.time() - t0 < 300


def test_no_device_memory_left_behind():
    """The drop-in loops allocate a storage context per snapshot and the reference's users drop the returned lists when
    they are done: device memory must come back (torch arenas) and repeated context creation must not leak driver-side
    objects (streams, events, scratch) either."""
    import gc
    import torch
    from time_crystal_tensor_network_b200.engine import Context
    model = KickedIsingModel(n_sites=10, J=1.0, h_disorder=0.3, tau=1.0, disorder_seed=3)
    model.pi_pulse_gate = __import__('scipy.linalg').linalg.expm(-1j * np.pi / 2 * 0.9 * model.sigma_x)

    def one_run():
        states, times, info = CustomFloquet(model, dict(chi_max=16)).evolve_floquet(create_initial_state(10, 'neel'), 12)
        return float(magnetization(states[-1], 'z'))

    first = one_run()
    gc.collect()
    torch.cuda.synchronize()
    base_alloc = torch.cuda.memory_allocated()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(5):
        assert one_run() == first                      # and the runs are reproducible bit for bit
    for _ in range(200):
        c = Context(12, 16, 2)
        c.set_product_state([[0, 1] * 6] * 2)
        c.close()
    gc.collect()
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() == base_alloc
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, f'{(free0 - free1) >> 20} MiB of device memory gone after 200 contexts'


def test_library_owned_memory_without_torch():
    """A process that never imports PyTorch (``python main.py``, the reference's scripts on the drop-in modules) runs on
    arenas and streams owned by the library; the records have the bits of the torch-owned run, and torch stays out."""
    import json
    import subprocess
    code = '''
import sys, json, time
t0 = time.perf_counter()
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import numpy as np, scipy.linalg
from core.tensor_utils import create_initial_state
from core.observables import magnetization, calculate_loschmidt_echo
from models.kicked_ising import KickedIsingModel
from dynamics.tebd_evolution import CustomFloquet
model = KickedIsingModel(n_sites=10, J=1.0, h_disorder=0.3, tau=1.0, disorder_seed=3)
model.pi_pulse_gate = scipy.linalg.expm(-1j * np.pi / 2 * 0.9 * model.sigma_x)
psi0 = create_initial_state(10, 'neel')
states, times, info = CustomFloquet(model, dict(chi_max=16)).evolve_floquet(psi0, 12)
out = {'z': [magnetization(s, 'z', site=3) for s in states], 'le': [calculate_loschmidt_echo(psi0, s) for s in states],
       'chi': [int(max(s.chi)) for s in states], 'torch': 'torch' in sys.modules, 'seconds': time.perf_counter() - t0}
print(json.dumps(out))
''' % (os.path.join(ROOT, 'src'), ROOT)
    runs = {}
    for mode in ('', 'torch'):
        env = dict(os.environ, TC_ARENA=mode)
        res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        runs[mode] = json.loads(res.stdout.strip().splitlines()[-1])
    assert runs['']['torch'] is False and runs['torch']['torch'] is True
    for key in ('z', 'le', 'chi'):
        assert runs[''][key] == runs['torch'][key], key
    print('process wall time without / with torch: %.2f / %.2f s' % (runs['']['seconds'], runs['torch']['seconds']))
