"""CPU suite, part 2: host-side logic of the product (no GPU): spectral post-processing against the
golden vectors, the C-ABI library's symbol table against include/tc_b200.h, loud failure without CUDA,
sharding helpers (incl. a world_size-2 gloo run), bench bookkeeping."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spectral_postprocessing_matches_reference(golden):
    from time_crystal_tensor_network_b200.core import observables as obs
    for key, ent in golden['post'].items():
        name, T = key.split('|T=')
        period = float(T)
        s = np.array(golden['series'][name])
        t = np.arange(len(s)) * period
        fund, sub = obs.subharmonic_response(list(s), period)
        assert [float(fund), float(sub)] == ent['subharmonic_response']
        assert obs.extract_subharmonic_amplitude(t, s, period) == ent['extract_subharmonic_amplitude']
        assert obs.extract_subharmonic_amplitude_from_loschmidt(t, s, period) == ent['extract_from_loschmidt']
        assert float(obs.detect_period_doubling_from_loschmidt(list(np.abs(s)))) == ent['detect_period_doubling']
    assert float(obs.fidelity_decay(list(np.exp(-0.05 * np.arange(20) * 2.0)), list(np.arange(20) * 2.0))) == \
        golden['fidelity_decay']
    # 31-sample +-1 series, period 2 -> positive-frequency bin 14 of 15, amplitude 1.0
    k = np.arange(31)
    assert obs.extract_subharmonic_amplitude(k * 2.0, (-1.0) ** k, 2.0) == 1.0
    assert obs.extract_subharmonic_amplitude(np.arange(5.0), np.ones(5), 2.0) == 0.0


def test_pauli_and_unused_gate_helper(golden):
    from time_crystal_tensor_network_b200.core.tensor_utils import pauli_matrices, create_time_evolution_gates
    p = pauli_matrices()
    for k, v in golden['pauli'].items():
        assert np.array_equal(p[k], np.array(v['re']) + 1j * np.array(v['im']))
    assert np.allclose(p['X'] @ p['Y'], 1j * p['Z'])
    g = create_time_evolution_gates(1.0, 0.3, 0.5, 4)
    # literal reproduction of the reference's element-wise exp: off-diagonal entries are exp(0) = 1
    assert g['ising_evolution'][0, 1] == 1.0 and g['pi_pulse'][0, 0] == 1.0


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'tc_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tc_[A-Za-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from time_crystal_tensor_network_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/tc_b200.h but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes prototype'
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.tc_version() >= 100
    assert lib.tc_ctx_arena_bytes(32, 128, 32) > 1.5 * 1024 ** 3     # 0.5 GiB of state + 1 GiB of workspaces
    assert lib.tc_ctx_arena_bytes(0, 1, 1) == 0
    # a storage-only context (snapshots) carries no SVD workspace: the 32-chain metric shape needs a third of the bytes,
    # a single L = 64, chi = 256 snapshot 0.27 GB instead of 0.8 GB
    assert lib.tc_ctx_arena_bytes2(32, 128, 32, 0) == lib.tc_ctx_arena_bytes(32, 128, 32)
    assert lib.tc_ctx_arena_bytes2(32, 128, 32, 1) < 0.6 * 1024 ** 3
    assert lib.tc_ctx_arena_bytes2(64, 256, 1, 1) < 0.4 * lib.tc_ctx_arena_bytes2(64, 256, 1, 0)


def test_documented_import_routes():
    """Every import route INTEGRATION.md documents works from a clean interpreter (no GPU needed to import)."""
    import subprocess
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    assert "/path/to/repo/src'" in doc and "/path/to/repo/time_crystal_tensor_network_b200'" not in doc
    for code in (
        "import sys; sys.path.insert(0, %r)\n"
        "from core.tensor_utils import create_initial_state\n"
        "from core.observables import calculate_loschmidt_echo, magnetization, staggered_magnetization\n"
        "from models.kicked_ising import KickedIsingModel\n"
        "from dynamics.tebd_evolution import CustomFloquet, TEBDEvolution\n" % os.path.join(ROOT, 'src'),
        "import sys; sys.path.insert(0, %r)\n"
        "from time_crystal_tensor_network_b200.models.kicked_ising import KickedIsingModel\n"
        "from time_crystal_tensor_network_b200.dynamics.tebd_evolution import CustomFloquet\n"
        "from time_crystal_tensor_network_b200.core.observables import magnetization\n" % ROOT,
    ):
        out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/')
        assert out.returncode == 0, out.stderr


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is visible')
    from time_crystal_tensor_network_b200.engine import Context, EngineError
    from time_crystal_tensor_network_b200.core.tensor_utils import create_initial_state
    with pytest.raises(EngineError):
        Context(4, 2, 1)
    with pytest.raises(EngineError):
        create_initial_state(4, 'neel')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'time_crystal_tensor_network_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                assert 'oracle' not in open(os.path.join(base, f)).read(), f'{f} mentions the oracle'


def test_only_tests_smoke_and_bench_import_the_oracle():
    """oracle/ is test infrastructure: besides tests/ only __graft_entry__.smoke() and bench.py's CPU legs import it --
    not main.py, not the drop-in modules under src/, not the scripts."""
    import re
    pat = re.compile(r'^\s*(from|import)\s+oracle\b', re.M)
    offenders = []
    for base in (ROOT, os.path.join(ROOT, 'scripts'), os.path.join(ROOT, 'src')):
        for dirpath, dirs, files in os.walk(base):
            if base == ROOT:
                dirs[:] = []                       # top level only here; tests/ and oracle/ are excluded by construction
            for f in files:
                if f.endswith('.py') and pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert sorted(offenders) == ['__graft_entry__.py', 'bench.py'], offenders


def test_no_cpu_fallback_on_library_owned_memory():
    """The torch-free route (arena and stream owned by the library) fails as loudly without a GPU as the torch one."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from time_crystal_tensor_network_b200.engine import Context, EngineError\n"
            "try:\n    Context(4, 2, 1)\nexcept EngineError as e:\n    print('EngineError')\n"
            "print('torch' in sys.modules)\n" % ROOT)
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is visible')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/', env=dict(os.environ, TC_ARENA=''))
    assert out.returncode == 0 and out.stdout.split() == ['EngineError', 'False'], out.stdout + out.stderr


def test_shard_bounds():
    from time_crystal_tensor_network_b200.sharding import shard_bounds, shard_sizes
    for n in (0, 1, 7, 32, 256, 1024, 1025):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
            sizes = shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    assert shard_bounds(256, 8, 3) == (96, 128)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from time_crystal_tensor_network_b200.sharding import shard_bounds, gather_records, disorder_average
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    n, T, L = 5, 3, 4                                    # ragged: shards of 3 and 2 chains
    full = np.arange(T * n * L, dtype=float).reshape(T, n, L)
    lo, hi = shard_bounds(n, world, rank)
    got = gather_records(full[:, lo:hi], n, axis=1)
    cfull = full + 1j * full[::-1]
    cgot = gather_records(cfull[:, lo:hi], n, axis=1)
    avg = disorder_average(full[:, lo:hi].sum(axis=1), hi - lo)
    q.put((rank, np.array_equal(got, full), np.array_equal(cgot, cfull), np.allclose(avg, full.mean(axis=1))))
    dist.destroy_process_group()


def test_gather_world_size_two_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res)


def test_bench_flop_model():
    sys.path.insert(0, ROOT)
    import bench
    L, chi = 32, 128
    ft, fs, fb = bench.update_flops(np.full(L + 1, chi))
    # uniform-chi count of SURVEY 8d: (32 + 672 + 32) chi^3 per update, 2(L-1) updates per period
    n = 2 * (L - 1)
    assert ft == n * 32 * chi ** 3 and fb == n * 32 * chi ** 3 and fs == n * 672 * chi ** 3


def test_bench_arms_share_one_config():
    """Both arms of bench.py describe their workload with the same function; weak and strong chain counts."""
    sys.path.insert(0, ROOT)
    import bench
    import argparse
    ns = argparse.Namespace(L=32, chi=128, eps=0.1, prep_eps=0.3, prep_periods=10, chains=32, scaling='weak', total_chains=256)
    assert bench.chains_of(ns, 8) == (32, 256)
    cfg = bench.make_config(ns, 8)
    assert cfg['chains_total'] == 256 and cfg['chains_per_gpu'] == 32 and cfg['seeds'] == '1000..1255'
    ns.scaling = 'strong'
    assert bench.chains_of(ns, 8) == (32, 256) and bench.chains_of(ns, 3) == (86, 256) and bench.chains_of(ns, 1) == (256, 256)
    src = open(os.path.join(ROOT, 'bench.py')).read()
    assert src.count("'config': make_config(a, world)") == 1 and src.count('cfg = make_config(a, world)') == 1   # reference arm, GPU arm


def test_reference_import_layout():
    """With only <repo>/src on sys.path the reference's import lines work unchanged (main.py:32-37)."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from core.tensor_utils import create_initial_state, pauli_matrices\n"
        "from core.observables import calculate_loschmidt_echo, staggered_magnetization, magnetization\n"
        "from core import create_initial_state as c2\n"
        "from models.kicked_ising import KickedIsingModel\n"
        "from models import KickedIsingModel as K2\n"
        "from dynamics.tebd_evolution import TEBDEvolution, CustomFloquet\n"
        "from dynamics import TEBDEvolution as T2\n"
        "m = KickedIsingModel(4, 1.0, 0.2, 1.0, disorder_seed=42)\n"
        "print(repr(float(m.h_fields[0])), len(m.ising_gates), m.pi_pulse_gate[0, 1])\n"
    ) % os.path.join(ROOT, 'src')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/')
    assert out.returncode == 0, out.stderr
    assert abs(float(out.stdout.split()[0]) + 0.050183952461055) < 1e-16 and out.stdout.split()[1] == '3'


def test_main_drivers_host_side(golden, tmp_path):
    """main.py: config parser on the repo's config.txt and on a temp file, DTC detector and Fourier spectrum
    against the values the reference's own functions produced (bit-exact)."""
    sys.path.insert(0, ROOT)
    import main as m
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        params = m.read_parameters('config.txt')
    finally:
        os.chdir(cwd)
    assert params == golden['config_params']
    f = tmp_path / 'p.txt'
    f.write_text('# c\nA = 3\nB = 0.5  # x\nC = [1, 2, 3]\nD = [0.5, 1]\nE = neel\nF = a,b\nG = 1,2\nH = []\nK = 1e-3\n\nbad line\n')
    p = m.read_parameters(str(f))
    assert p == {'A': 3, 'B': 0.5, 'C': [1, 2, 3], 'D': [0.5, 1.0], 'E': 'neel', 'F': ['a', 'b'], 'G': [1, 2], 'H': [],
                 'K': 0.001}
    os.chdir(str(tmp_path))
    try:
        assert m.read_parameters('does_not_exist.txt') == {}
    finally:
        os.chdir(cwd)
    for key, ent in golden['post'].items():
        name, T = key.split('|T=')
        period = float(T)
        s = np.array(golden['series'][name])
        t = np.arange(len(s)) * period
        with np.errstate(invalid='ignore', divide='ignore'):
            assert float(m.stringent_dtc_detection(list(np.abs(s)), list(t), period)) == ent['stringent_dtc_detection']
        if len(s) > 2:
            fr, pw = m.calculate_fourier_spectrum(t, s, period)
            assert np.array_equal(fr, ent['fourier_freqs']) and np.array_equal(pw, ent['fourier_power'])
            assert int(np.argmax(pw)) == ent['fourier_peak_bin']
    assert m.stringent_dtc_detection([1.0] * 10, list(range(10)), 2.0) == 0.0
    bad = m.calculate_phase_point(0.2, 2.0, {})          # missing keys -> swallowed, success=False
    assert bad['success'] is False and bad['A2T'] == 0.0 and bad['avg_bond_dim'] == 1.0
