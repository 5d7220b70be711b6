"""Parity of the CUDA path against the LAPACK TEBD oracle at the sizes the benchmark and the BASELINE configurations
are quoted on (north star: <Z_i>(t), bond entropies, Loschmidt echo <= 1e-8 absolute at equal chi_max and truncation
cut-offs; spectral peak positions exactly).

The oracle (oracle/tebd_ref.py, update_bond_tebd + truncate(): /root/reference/src/models/kicked_ising.py:100-126,162-188
on TeNPy's TEBD update) needs tens of seconds per chain at these sizes, so its records are committed as fixtures
(tests/golden/headline_*.npz etc., written by oracle/make_headline_golden.py) and re-computed live with TC_GOLDEN_LIVE=1.

Every comparison goes through the C ABI (FloquetEnsemble -> tc_floquet_run_host).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import tebd_ref  # noqa: E402  (checker only)
from oracle.make_headline_golden import CASES  # noqa: E402

TOL = 1e-8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle(name):
    """{key_k: array} of the oracle for every seed of the case: the committed fixture, or a live run."""
    c = CASES[name]
    if os.environ.get('TC_GOLDEN_LIVE') == '1':
        out = {}
        for k, seed in enumerate(c['seeds']):
            h = tebd_ref.disorder_fields(c['L'], c['W'], seed)
            r = tebd_ref.run_schedule(c['L'], 1.0, h, 1.0, c['schedule'], mode='tebd', trunc=c['trunc'])
            for key in ('Z', 'S_ent', 'LE', 'chi'):
                out[f'{key}_{k}'] = r[key]
            out[f'h_{k}'] = h
        return out
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')))


def _gpu_schedule(name, env=None):
    """The case's chains in one context, the kick imperfection switched between the phases of the schedule (the Ising
    gates do not depend on it): records before the first period and after every period."""
    from time_crystal_tensor_network_b200 import engine as eng
    c = CASES[name]
    L = c['L']
    hs = np.array([eng.disorder_fields(L, c['W'], s) for s in c['seeds']])
    ens = eng.FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=c['schedule'][0][0], mode='tebd', state='neel', **c['trunc'])
    parts = []
    for eps, n in c['schedule']:
        kick = np.ascontiguousarray(np.broadcast_to(eng.kick_matrix(eps), (len(hs), 2, 2)))
        ens.ctx.set_model(ens.gates, kick)
        parts.append(ens.run(n))
    out = {k: np.concatenate([p[k] for p in parts], axis=0) for k in ('Z', 'S_ent', 'LE', 'chi')}
    out['flags'] = parts[-1]['flags']
    out['hs'] = hs
    ens.close()
    return out


def _compare(name, out, ref, tol=TOL):
    c = CASES[name]
    worst = {}
    for k in range(len(c['seeds'])):
        assert np.array_equal(out['hs'][k], ref[f'h_{k}'])
        chi_ref = ref[f'chi_{k}']
        assert np.array_equal(out['chi'][:, k, 1:-1], chi_ref), \
            f'{name} chain {k}: bond-dimension tables differ first at record ' \
            f'{int(np.argmax(np.any(out["chi"][:, k, 1:-1] != chi_ref, axis=1)))}'
        for key in ('Z', 'S_ent', 'LE'):
            dev = np.abs(out[key][:, k] - ref[f'{key}_{k}'])
            per_t = dev.reshape(dev.shape[0], -1).max(axis=1)
            worst[(key, k)] = per_t
    msg = '; '.join(f'{key}[{k}] max {v.max():.2e} at record {int(v.argmax())}' for (key, k), v in worst.items())
    print(f'{name}: {msg}')
    for (key, k), v in worst.items():
        assert v.max() < tol, f'{name}: {msg}'
    assert out['flags']['svd_not_converged'] == 0 and out['flags']['chi_cap_overflow'] == 0


def test_headline_shape_against_oracle(engine):
    """bench.py's workload: L = 32, chi_max = 128, svd_min 1e-12, trunc_cut 1e-7, two chains (seeds 1000, 1001), nine
    periods at eps = 0.3 (the central bonds reach chi_max in period 8: theta = 256 x 256 with the chi_max cut
    active, the `sweeps<8, true>` instance of the blocked Jacobi kernel) and three periods at eps = 0.1."""
    name = 'headline_L32_chi128'
    out = _gpu_schedule(name)
    assert out['chi'].max() == 128 and out['chi'][-1, :, 16].min() == 128
    _compare(name, out, _oracle(name))


@pytest.mark.parametrize('cluster', ['auto', '1', '2', '8', 'wide_v1'])
def test_wide_matrices_truncated_against_oracle(engine, cluster, monkeypatch):
    """chi_max = 256 binding (theta 512 x 512: the wide QR instance and the team Jacobi kernel on a thread-block
    cluster, BASELINE config 4's regime) at L = 18, where an untruncated bond would reach 512.  The cluster size
    (TC_WIDE_CLUSTER: CTAs per matrix; automatic = as many as fill the GPU) changes the pair order, not the result;
    wide_v1 is the warp-per-pair cluster kernel of round 1."""
    if cluster == 'wide_v1':
        monkeypatch.setenv('TC_JACOBI', 'wide_v1')
    elif cluster != 'auto':
        monkeypatch.setenv('TC_WIDE_CLUSTER', cluster)
    name = 'wide_L18_chi256'
    out = _gpu_schedule(name)
    assert out['chi'].max() == 256
    _compare(name, out, _oracle(name))


@pytest.mark.parametrize('case', ['wide_L18_chi256', 'headline_L32_chi128'])
def test_qr_on_a_cluster_is_bit_identical(engine, case, monkeypatch):
    """The blocked QR kernel on a thread-block cluster (TC_QR_CLUSTER CTAs per matrix; automatic for launches with few
    matrices): every CTA factorises the panel with the same arithmetic and the trailing tiles are only dealt out
    differently, so the triangular factor -- and with it every record of the run -- has the same bits for any cluster
    size.  Both QR instances (256 and 512 rows), truncation active."""
    name = case
    monkeypatch.setenv('TC_GROUPS', '1')
    outs = {}
    for cs in ('1', '2', '4', '8'):
        monkeypatch.setenv('TC_QR_CLUSTER', cs)
        outs[cs] = _gpu_schedule(name)
    for cs in ('2', '4', '8'):
        for key in ('Z', 'S_ent', 'LE', 'chi'):
            assert np.array_equal(outs['1'][key], outs[cs][key]), f'{name}: {key} differs between QR cluster 1 and {cs}'


@pytest.mark.parametrize('small_kernel', ['0', '2', '3', '4'])
def test_phase_diagram_point_shape_against_oracle(engine, small_kernel, monkeypatch):
    """BASELINE config 3's grid point: L = 24, chi_max = 64, 20 periods at eps = 0.1, with every Jacobi kernel a narrow
    context can run: the half-warp-row kernel in its 8-warp instance (the default for ensembles with more matrices per
    layer than SMs, forced here for the single chain with TC_SMALL_KERNEL=3) and its 16-warp instance (=4, what a
    single chain runs by default), the full-warp 16-warp kernel (=0) and its instance with row blocks of 8 (=2)."""
    monkeypatch.setenv('TC_SMALL_KERNEL', small_kernel)
    name = 'c3_L24_chi64'
    out = _gpu_schedule(name)
    _compare(name, out, _oracle(name))


def test_subharmonic_peak_position_of_an_entangling_run(engine):
    """Period-doubling spectral peak positions exactly: 64 periods at eps = 0.1 (bond dimension up to 32, truncation
    active); the FFT bin of the largest positive-frequency component of the staggered magnetisation and of the
    Loschmidt echo are the oracle's, and the reference's own extractors (src/core/observables.py:153-221,372-439,
    main.calculate_fourier_spectrum) return the same numbers on both series."""
    import sys
    sys.path.insert(0, ROOT)
    import main as m
    from time_crystal_tensor_network_b200.core import observables as obs
    name = 'dtc_L12_chi32'
    c = CASES[name]
    out = _gpu_schedule(name)
    ref = _oracle(name)
    _compare(name, out, ref)
    L, T = c['L'], 2.0
    sign = (-1.0) ** np.arange(L)
    times = np.arange(out['Z'].shape[0]) * T
    for k in range(len(c['seeds'])):
        stag = (out['Z'][:, k] * sign).mean(axis=1)
        stag_ref = (ref[f'Z_{k}'] * sign).mean(axis=1)

        def peak_bin(series):
            w = (series - series.mean()) * np.hanning(len(series))
            f = np.fft.fft(w)
            fr = np.fft.fftfreq(len(w), d=T)
            pos = np.nonzero(fr > 0)[0]
            return int(pos[np.argmax(np.abs(f[pos]))]), fr

        b, fr = peak_bin(stag)
        b_ref, _ = peak_bin(stag_ref)
        assert b == b_ref
        assert abs(fr[b] - 0.5 / T) <= 1.01 / (len(times) * T)           # the sub-harmonic bin (half the drive frequency)
        b_le, _ = peak_bin(out['LE'][:, k])
        b_le_ref, _ = peak_bin(ref[f'LE_{k}'])
        assert b_le == b_le_ref
        a = obs.extract_subharmonic_amplitude(times, stag, T)
        a_ref = obs.extract_subharmonic_amplitude(times, stag_ref, T)
        assert abs(a - a_ref) < TOL
        f1, p1 = m.calculate_fourier_spectrum(times, stag, T)
        f2, p2 = m.calculate_fourier_spectrum(times, stag_ref, T)
        assert int(np.argmax(p1)) == int(np.argmax(p2)) and np.array_equal(f1, f2)
        assert abs(obs.extract_subharmonic_amplitude_from_loschmidt(times, out['LE'][:, k], T) -
                   obs.extract_subharmonic_amplitude_from_loschmidt(times, ref[f'LE_{k}'], T)) < 1e-6


@pytest.mark.parametrize('switch', ['TC_THRESH=0', 'TC_EARLY_STOP=0'])
def test_svd_shortcuts_change_nothing_at_rounding_level(engine, switch, monkeypatch):
    """The pattern DESIGN.md prescribes for every shortcut inside the SVD: same saturated state (256 x 256 theta with
    the chi_max cut active), one more period with the shortcut on and off, agreement at rounding level (1e-12), far
    below what the 1e-8 oracle comparisons could see.  Shortcuts: threshold sweeps (TC_THRESH=0 rotates every pair in
    every sweep), the quadratic-convergence stopping rule (TC_EARLY_STOP=0 runs the verification sweep)."""
    from time_crystal_tensor_network_b200 import engine as eng
    monkeypatch.setenv('TC_ARENA', 'torch')                 # the state is cloned arena to arena below
    L, chi = 32, 128
    hs = np.array([eng.disorder_fields(L, 0.3, 1000)])
    kw = dict(epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
    base = eng.FloquetEnsemble(L, 1.0, 1.0, hs, **kw)
    base.run(9)
    assert base.ctx.chi()[0, L // 2] == chi
    key, val = switch.split('=')
    monkeypatch.setenv(key, val)
    other = eng.FloquetEnsemble(L, 1.0, 1.0, hs, **kw)       # the switch is read when a context is created
    monkeypatch.delenv(key)
    other.ctx._arena.copy_(base.ctx._arena)
    a = base.run(1, measure_now=False)
    b = other.run(1, measure_now=False)
    assert np.array_equal(a['chi'], b['chi'])
    for k in ('Z', 'S_ent', 'LE'):
        assert np.max(np.abs(a[k] - b[k])) < 1e-12, k
    # the kept Schmidt values themselves, relative
    for bond in (L // 2 - 1, L // 2, L // 2 + 1):
        sa, sb = base.ctx.get_S(0, bond), other.ctx.get_S(0, bond)
        assert np.max(np.abs(sa - sb) / sa) < 1e-10
    base.close()
    other.close()


@pytest.mark.parametrize('L,chi,prep,variants', [
    (14, 64, 8, ['TC_SMALL_KERNEL=3', 'TC_SMALL_KERNEL=4', 'TC_SMALL_KERNEL=2', 'TC_SMALL_KERNEL=0']),   # 128 columns: half-warp rows (8 / 16 warps), full-warp rows in blocks of 8 / 16
    (16, 128, 9, ['TC_JACOBI=blocked', 'TC_JACOBI=team']),
    (18, 256, 10, ['TC_WIDE_CLUSTER=1', 'TC_WIDE_CLUSTER=4', 'TC_WIDE_CLUSTER=8', 'TC_JACOBI=wide_v1']),
])
def test_svd_kernels_repeatedly_against_lapack(engine, L, chi, prep, variants, monkeypatch):
    """Stress test of the batched SVD stage on its own: both parity layers applied 12 times from the same saturated
    state with every Jacobi kernel variant; the singular values the kernel leaves (row norms of the workspace) must be
    LAPACK's of theta = diag(S_i x 1_2) C every time, to 1e-12 of the largest.  (A race between the two warps of a
    team left half a row unscaled in about one run in five of the first version of the team kernel; the 1e-8
    comparisons on observables see that only when it hits a well-populated Schmidt value.)"""
    from time_crystal_tensor_network_b200 import engine as eng
    from time_crystal_tensor_network_b200 import _lib
    monkeypatch.setenv('TC_GROUPS', '1')
    monkeypatch.setenv('TC_ARENA', 'torch')                 # the state is cloned arena to arena below
    hs = np.array([eng.disorder_fields(L, 0.3, 11)])
    kw = dict(epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
    base = eng.FloquetEnsemble(L, 1.0, 1.0, hs, **kw)
    base.ctx.floquet_step(prep)
    base.ctx.sync()
    chi0 = base.ctx.chi()[0]
    assert chi0.max() == chi
    for var in variants:
        key, val = var.split('=')
        monkeypatch.setenv(key, val)
        e = eng.FloquetEnsemble(L, 1.0, 1.0, hs, **kw)
        monkeypatch.delenv(key)
        for parity in (0, 1):
            sv = {}
            for rep in range(12):
                e.ctx._arena.copy_(base.ctx._arena)
                e.ctx.apply_layer(parity, 0)
                e.ctx.sync()
                for jb in range(L // 2):
                    i = 2 * jb + parity
                    if i + 1 >= L:
                        continue
                    M, N = 2 * int(chi0[i]), 2 * int(chi0[i + 2])
                    if min(M, N) < 64:
                        continue
                    if rep == 0:
                        C = e.ctx.dbg_get(_lib.DBG_C, 0, jb, (M, N), np.complex128)
                        sv[jb] = np.linalg.svd(C * np.repeat(base.ctx.get_S(0, i), 2)[:, None], compute_uv=False)
                    w = np.sort(e.ctx.dbg_get(_lib.DBG_W, 0, jb, (min(M, N),), np.float64))[::-1]
                    err = np.max(np.abs(w - sv[jb])) / sv[jb][0]
                    assert err < 1e-12, (var, parity, rep, jb, M, N, err)
        assert e.ctx.flags()['svd_not_converged'] == 0
        e.close()
    base.close()
