"""Randomised differential test of the CUDA path (through the C ABI) against the LAPACK TEBD oracle: chain length, coupling,
period, disorder, kick imperfection, truncation rule, bond-dimension caps that are not powers of two (ragged row blocks
and 8-column tiles that straddle cache lines in the QR / Jacobi kernels), several chains per context.

The regular GPU suite runs TC_FUZZ_CASES = 16 cases (seconds); a soak is ``TC_FUZZ_CASES=300 python -m pytest
tests/test_gpu_fuzz.py -m gpu -q`` (profiles/README.md records the last one).  Every case is reproducible from its index.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import tebd_ref  # noqa: E402  (checker only)

TOL = 1e-8
N_CASES = int(os.environ.get('TC_FUZZ_CASES', '16'))
FIRST = int(os.environ.get('TC_FUZZ_FIRST', '0'))


def _case(k):
    rng = np.random.default_rng(77000 + k)
    mode = 'reference' if k % 4 == 3 else 'tebd'
    L = int(rng.integers(2, 17 if mode == 'tebd' else 11))
    c = dict(k=k, mode=mode, L=L, J=float(rng.uniform(0.5, 1.5)), tau=float(rng.uniform(0.3, 1.2)),
             eps=float(rng.choice([0.0, rng.uniform(0.02, 0.4)])), W=float(rng.uniform(0.0, 1.0)),
             n=int(rng.integers(3, 11)), state=str(rng.choice(['neel', 'all_up', 'all_down'])), R=int(rng.integers(1, 4)))
    if mode == 'tebd':
        c['trunc'] = dict(chi_max=int(rng.choice([1, 2, 3, 5, 6, 7, 9, 11, 13, 17, 20, 24, 33, 40, 48, 64])),
                          svd_min=float(rng.choice([1e-12, 1e-10])), trunc_cut=float(rng.choice([1e-10, 1e-7])))
    return c


@pytest.mark.parametrize('k', list(range(FIRST, FIRST + N_CASES)))
def test_random_case_against_tebd_oracle(engine, k):
    from time_crystal_tensor_network_b200.engine import FloquetEnsemble
    c = _case(k)
    L, R = c['L'], c['R']
    hs = np.array([tebd_ref.disorder_fields(L, c['W'], 500 + 10 * k + r) for r in range(R)])
    if c['mode'] == 'tebd':
        kw = dict(mode='tebd', **c['trunc'])
    else:
        kw = dict(mode='reference', chi_max=2 ** (L // 2))
    ens = FloquetEnsemble(L, c['J'], c['tau'], hs, epsilon=c['eps'], state=c['state'], **kw)
    out = ens.run(c['n'])
    flags = out['flags']
    ens.close()
    assert flags['svd_not_converged'] == 0 and flags['chi_cap_overflow'] == 0, c
    for r in range(R):
        ref = tebd_ref.run(L, c['J'], hs[r], c['tau'], c['n'], epsilon=c['eps'], state=c['state'], mode=c['mode'],
                           trunc=c.get('trunc'))
        if L > 1:
            assert np.array_equal(out['chi'][:, r, 1:-1], ref['chi']), (c, r)
            assert np.max(np.abs(out['S_ent'][:, r] - ref['S_ent'])) < TOL, (c, r)
        assert np.max(np.abs(out['Z'][:, r] - ref['Z'])) < TOL, (c, r)
        assert np.max(np.abs(out['LE'][:, r] - ref['LE'])) < TOL, (c, r)


N_WIDE = int(os.environ.get('TC_FUZZ_WIDE', '2'))


@pytest.mark.parametrize('k', list(range(N_WIDE)))
def test_random_wide_case_against_tebd_oracle(engine, k):
    """Bond-dimension caps between 129 and 256, odd ones included: theta up to 512 columns, i.e. the 512-row QR instance and
    the team Jacobi kernel on thread-block clusters, with ragged row blocks; strong kicks so that the cap binds."""
    from time_crystal_tensor_network_b200.engine import FloquetEnsemble
    rng = np.random.default_rng(99000 + k)
    L = int(rng.choice([16, 17]))
    R = int(rng.integers(1, 3))
    trunc = dict(chi_max=int(rng.integers(129, min(256, 2 ** (L // 2)) + 1)), svd_min=1e-12,
                 trunc_cut=float(rng.choice([1e-10, 1e-7])))
    eps, W, n = float(rng.uniform(0.25, 0.4)), float(rng.uniform(0.1, 0.8)), int(rng.integers(9, 12))
    hs = np.array([tebd_ref.disorder_fields(L, W, 900 + 10 * k + r) for r in range(R)])
    ens = FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=eps, mode='tebd', **trunc)
    out = ens.run(n)
    ens.close()
    c = dict(k=k, L=L, R=R, eps=eps, W=W, n=n, **trunc)
    assert out['flags']['svd_not_converged'] == 0 and out['flags']['chi_cap_overflow'] == 0, c
    assert out['chi'].max() > 128, c
    for r in range(R):
        ref = tebd_ref.run(L, 1.0, hs[r], 1.0, n, epsilon=eps, mode='tebd', trunc=trunc)
        assert np.array_equal(out['chi'][:, r, 1:-1], ref['chi']), (c, r)
        for key in ('Z', 'S_ent', 'LE'):
            assert np.max(np.abs(out[key][:, r] - ref[key])) < TOL, (c, r, key)
